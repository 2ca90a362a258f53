"""Synthetic geometry / weights / batches shared by the oracle, the tests and the benchmark.

The generators live in `valor_b200.synthetic` (the benchmark's GPU arm builds its inputs from them and may not
import anything under oracle/); this module re-exports them for the checker side."""
from valor_b200.synthetic import *  # noqa: F401,F403
from valor_b200.synthetic import BASE, TINY, Geometry, make_batch, make_state_dict, relative_position_index, token_masker  # noqa: F401
