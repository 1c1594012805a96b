#!/usr/bin/env python
"""Bring-up diagnostics for the GEMM kernels on a real B200: each (backend, a_mn, b_mn) case runs in its own
subprocess under a timeout, so a hang or a sticky CUDA error in one variant cannot hide the others.
Usage: python tools/gpu_diag.py            (driver)   |   python tools/gpu_diag.py case <backend> <a_mn> <b_mn>"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case(backend, a_mn, b_mn):
    import torch

    os.environ["MUSE_B200_GEMM"] = backend
    from open_muse_b200 import ops

    dev = "cuda"
    for (M, N, K) in [(128, 128, 64), (128, 128, 256), (128, 256, 64), (256, 256, 512), (1024, 1536, 512), (65792 // 8, 512, 2048)]:
        g = torch.Generator().manual_seed(M + N + K)
        A = torch.randn((K, M) if a_mn else (M, K), generator=g).to(torch.bfloat16).to(dev)
        B = torch.randn((K, N) if b_mn else (N, K), generator=g).to(torch.bfloat16).to(dev)
        ref = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t())
        C = torch.zeros(M, N, dtype=torch.float32, device=dev)
        ops.gemm(A, B, C, M, N, K, A.stride(0), B.stride(0), N, a_mn, b_mn, ops.EPI_F32)
        torch.cuda.synchronize()
        err = float((C - ref).norm() / ref.norm())
        bad = int(((C - ref).abs() > 1e-2 * ref.abs().max()).sum())
        print(f"  {backend} a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K}: rel-L2 {err:.3e}, bad elems {bad}/{M*N}", flush=True)
        if err > 1e-3 and M <= 256:
            # error structure: which 8-row / 16-col blocks are wrong
            e = ((C - ref).abs() > 1e-2 * ref.abs().max())
            rows = e.any(dim=1).nonzero().flatten().tolist()
            cols = e.any(dim=0).nonzero().flatten().tolist()
            print(f"    wrong rows {rows[:16]}... ({len(rows)}), wrong cols {cols[:16]}... ({len(cols)})")
            print("    C[0,:8]  ", C[0, :8].tolist())
            print("    ref[0,:8]", ref[0, :8].tolist())


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "case":
        return case(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    for backend in ("mma", "tcgen05"):
        for a_mn, b_mn in ((0, 0), (0, 1), (1, 1), (1, 0)):
            print(f"== {backend} a_mn={a_mn} b_mn={b_mn}", flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, "case", backend, str(a_mn), str(b_mn)], timeout=150,
                                   capture_output=True, text=True)
                print(r.stdout, end="")
                if r.returncode != 0:
                    print(f"  EXIT {r.returncode}\n{r.stderr[-1500:]}")
            except subprocess.TimeoutExpired as e:
                print(f"  TIMEOUT (hang?) partial output:\n{(e.stdout or b'').decode() if isinstance(e.stdout, bytes) else e.stdout}")


if __name__ == "__main__":
    main()
