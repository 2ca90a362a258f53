"""BASELINE configs[4] timing: 1 000 clips x 1 000 captions fine-grained similarity + recall on one B200 (features given,
as in test.py:validate_ret after the feature gather).  Prints one JSON line; the CPU leg times the reference formulation
(oracle.compute_fine_matrix: einsum + max/max, pretrain.py:178-211) on the host."""
import json, os, sys, time, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valor_b200 import functional as Fn, retrieval as R  # noqa: E402


def main():
    g = torch.Generator().manual_seed(5)
    Nt = Nv = 1000
    T, nV, nA, D = 32, 8, 2, 512
    ft = torch.nn.functional.normalize(torch.randn(Nt, T, D, generator=g), dim=-1)
    fb = torch.nn.functional.normalize(torch.randn(Nv, nV + nA, D, generator=g), dim=-1)
    lens = torch.randint(8, 31, (Nt,), generator=g)
    maskA = (torch.arange(T)[None, :] < lens[:, None]).to(torch.uint8)
    w_t, w_b = torch.randn(Nt, T, generator=g), torch.randn(Nv, nV + nA, generator=g)
    dev = {k: v.cuda() for k, v in dict(ft=ft.reshape(-1, D), fb=fb.reshape(-1, D), w_t=w_t, w_v=w_b[:, :nV].contiguous(),
                                        w_a=w_b[:, nV:].contiguous(), maskA=maskA).items()}
    ids = [f"v{i}" for i in range(Nv)]
    model = types.SimpleNamespace(contra_temp=types.SimpleNamespace(data=torch.tensor(0.07).cuda()))

    def step():
        sc = Fn.FineSimFn.apply(dev["ft"], dev["fb"], dev["w_t"], dev["w_v"], dev["w_a"], dev["maskA"], (Nt, Nv, T, nV, nA), ["tva"], True)[0]
        return R.compute_metric_ret(model, sc, ids, ids)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        log = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out = {"workload": "fast-retrieval-msrvtt similarity + recall: 1000 captions x 1000 clips, T=32, 8+2 slots, d=512 (BASELINE configs[4])",
           "gpu_ms": ms, "gemm_gflop": 2 * Nt * T * Nv * (nV + nA) * D / 1e9, "log": log}
    if "--cpu" in sys.argv:
        from oracle import valor_oracle as vo
        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        t0 = time.perf_counter()
        s = vo.compute_fine_matrix(ft, fb, maskA.long(), torch.ones(Nv, nV + nA, dtype=torch.long), w_t.clone(), w_b.clone())
        vo.compute_metric_ret(s, ids, ids)
        out["cpu_reference_formulation_s"] = time.perf_counter() - t0
        out["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
