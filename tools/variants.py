"""Build experiment variants of the library (extra -D flags on csrc/mlp_tc.cu etc.) into gaussianavatar_b200/variants/, and time them.
  python tools/variants.py build  name=-DGA_X=1,-DGA_Y=2 ...      (CPU box: nvcc)
  python tools/variants.py time   [S] [iters]                    (GPU box: decoder fwd+bwd per-kernel ms for every built variant)
Variants are throw-away measurement aids: nothing in the package loads them."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "gaussianavatar_b200", "variants")

def build(specs):
    from gaussianavatar_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    B.build_library()
    nvcc = B._nvcc()
    base_objs = sorted(glob.glob(os.path.join(B._OBJ, "*.o")))
    for spec in specs:
        name, _, flags = spec.partition("=")
        flags = [f for f in flags.split(",") if f]
        objs = []
        for o in base_objs:
            stem = os.path.basename(o)[:-2]
            if stem in ("mlp_tc", "raster", "decoder", "conv_tc"):
                vo = os.path.join(VDIR, f"{stem}_{name}.o")
                subprocess.run([nvcc, *B.NVCC_FLAGS, *flags, "-c", os.path.join(B._CSRC, stem + ".cu"), "-o", vo], check=True)
                objs.append(vo)
            else:
                objs.append(o)
        out = os.path.join(VDIR, f"lib_{name}.so")
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs, "-lcudart"], check=True)
        for o in objs:
            if o.startswith(VDIR): os.remove(o)
        print("built", out)

def time_one(path, S, iters):
    code = f"""
import sys, torch
sys.path.insert(0, {ROOT!r})
import gaussianavatar_b200._lib as L
L.LIB_PATH = {path!r}
from gaussianavatar_b200.network import POP_no_unet
torch.manual_seed(0)
net = POP_no_unet(c_geom=64, hsize=128).cuda()
geo = (torch.randn(1, 64, 128, 128) * 0.01).cuda().requires_grad_(True)
def step():
    dec = net.forward_packed(geo, {S}, 2); dec.backward(torch.ones_like(dec) * 1e-3)
for _ in range(2): step()
net.zero_grad(); geo.grad = None; step()
chk = "chk dgeo=%.6e dflat=%.6e" % (geo.grad.double().abs().sum().item(), sum(p.grad.double().abs().sum().item() for p in net.parameters() if p.grad is not None))
torch.cuda.synchronize(); L.profile(True)
for _ in range({iters}): step()
rep = L.profile_report()
print({os.path.basename(path)!r}, chk, " ".join(f"{{k}}={{v[1]/{iters}:.3f}}ms/{{v[0]//{iters}}}" for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1]) if k.startswith(("mlp_tc", "heads", "geom_conv", "sample_feat"))))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    print(r.stdout.strip() or ("FAILED " + os.path.basename(path) + "\n" + r.stderr[-800:]))

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
        iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
        for pth in sorted(glob.glob(os.path.join(VDIR, "lib_*.so"))):
            time_one(pth, S, iters)
