"""Full-size sanity / throughput run of a BASELINE.json config (not the headline bench line).

    python tools/run_config.py cfg3|cfg4|cfg1|cfg2 [precision] [batch]
"""
import sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import moleculediffusiontransformer_b200 as mdt
INV64 = dict(max_length=64, pred_dim=16, channels=64, unet_type="cfg", context_embedding_max_length=12,
             pos_emb_fourier=True, pos_emb_fourier_add=False, text_embed_dim=64, embed_dim_position=64)
FWD64 = dict(INV64, pred_dim=1, context_embedding_max_length=64)
WIDE = dict(INV64, max_length=128, pred_dim=32, channels=128)

name = sys.argv[1]
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
torch.manual_seed(0)
if name == "cfg3":
    model, B, n, cs, T = mdt.QMDiffusionForward(**FWD64).eval(), 16384, 64, 1.0, 64
elif name == "cfg4":
    model, B, n, cs, T = mdt.QMDiffusion(**WIDE).eval(), 8192, 12, 7.5, 128
elif name == "cfg1":
    model, B, n, cs, T = mdt.QMDiffusion(**INV64).eval(), 4, 12, 1.0, 64
else:
    model, B, n, cs, T = mdt.QMDiffusion(**INV64).eval(), 4096, 12, 7.5, 64
if len(sys.argv) > 3:
    B = int(sys.argv[3])
g = torch.Generator().manual_seed(2)
seq = torch.rand(B, n, generator=g) * 2 - 1
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = model.sample(seq, "cuda:0", cond_scale=cs, timesteps=T, seed=5, precision=prec)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name} {prec} B={B} T={T} cs={cs}: {dt*1e3:.1f} ms -> {B/dt:.1f} samples/s  finite={bool(torch.isfinite(out).all())} "
          f"mem={torch.cuda.mem_get_info()[0]/2**30:.1f} GiB free", flush=True)
