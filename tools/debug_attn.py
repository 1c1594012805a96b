import os, sys, math, subprocess
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run(B, nh, S, mode):
    os.environ["MUSE_B200_ATTN"] = mode
    from open_muse_b200 import ops
    H = nh * 64
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B * S, 3 * H, generator=g).to(torch.bfloat16).cuda()
    do = torch.randn(B * S, H, generator=g).to(torch.bfloat16).cuda()
    o, lse = ops.attn_fwd(qkv[:, :H], qkv[:, H:2*H], qkv[:, 2*H:], B, nh, S, S, 0.125)
    d = torch.full_like(qkv, float('nan'))
    ops.attn_bwd(qkv[:, :H], qkv[:, H:2*H], qkv[:, 2*H:], o, do, lse, d[:, :H], d[:, H:2*H], d[:, 2*H:], B, nh, S, S, 0.125)
    torch.cuda.synchronize()
    return o.float().cpu(), lse.cpu(), d.float().cpu()

if len(sys.argv) > 1:
    B, nh, S, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    torch.save(run(B, nh, S, mode), f"/tmp/attn_{mode}.pt")
    sys.exit(0)

for (B, nh, S) in [(2, 2, 257), (3, 2, 257), (4, 8, 257), (4, 8, 256), (3, 4, 385), (8, 8, 128)]:
    for mode in ("legacy", "tc"):
        subprocess.run([sys.executable, __file__, str(B), str(nh), str(S), mode], check=True)
    o1, l1, d1 = torch.load("/tmp/attn_legacy.pt"); o2, l2, d2 = torch.load("/tmp/attn_tc.pt")
    H = nh * 64
    def rel(a, b): return float((a - b).norm() / b.norm())
    msg = f"B={B} nh={nh} S={S}: o {rel(o2,o1):.2e} lse {rel(l2,l1):.2e}"
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2*H)), ("dv", slice(2*H, 3*H))):
        a, b = d2[:, sl].view(B, S, nh, 64), d1[:, sl].view(B, S, nh, 64)
        msg += f" | {name} {rel(a,b):.2e} nan={int(torch.isnan(a).sum())}"
        if rel(a, b) > 0.05 or torch.isnan(a).any():
            bad = ((a - b).abs() > 0.05 * b.abs().max()) | torch.isnan(a)
            idx = bad.any(-1).nonzero()
            msg += f" bad(b,s,h) first {idx[:3].tolist()} count {len(idx)} ; bad batches {sorted(set(idx[:,0].tolist()))} rows[min,max]=({int(idx[:,1].min())},{int(idx[:,1].max())})"
    print(msg, flush=True)
