"""Per-entry event timeline of four CTAs of the forward chain kernel (debug bit 128), config 2."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_gan_fd_b200 as b200
from sr_gan_fd_b200 import lib
L = lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=23).to(dev).eval()
lr = torch.rand(16, 3, 64, 64, device=dev)
NE, E0, NC = 256, 300, 4
with torch.no_grad():
    for _ in range(2): net(lr)
    L.b200sr_debug_set(128 | int(os.environ.get('DBG', 0)))
    net(lr); torch.cuda.synchronize()
    L.b200sr_debug_set(0)
n = 160 * 12 + NC * NE * 16
buf = (C.c_ulonglong * n)()
lib.check(L.b200sr_debug_read_profile(buf, n))
full = np.array(buf, dtype=np.float64)
print(f"sustained SM clock during the launch: {full[160*12-16] / full[160*12-15] * 1e3:.0f} MHz ({full[160*12-15]/1e6:.3f} ms)")
t = full[160 * 12:].reshape(NC, NE, 16)
t0 = t[t > 0].min()
t = np.where(t > 0, (t - t0) / 1e3, np.nan)  # us
names = ["prod@", "depOK", "mma0", "Aland", "mmaEnd", "accRdy", "stored", "signal", "epi@", "epiDep", "tmemLd", "sigArr", "preSt", "postSt", "pAfree", "pW0free"]
lo = int(os.environ.get("LO", 40)); hi = int(os.environ.get("HI", 64))
for c in range(0 if os.environ.get('BRIEF') else int(os.environ.get('NCP', NC))):
    print(f"--- CTA {c * 37}: entries {E0 + lo}..{E0 + hi - 1} (us)")
    print("entry " + " ".join(f"{x:>8s}" for x in names))
    for e in range(lo, hi):
        print(f"{E0 + e:5d} " + " ".join(f"{x:8.2f}" for x in t[c, e]))
d = t[0, lo:hi, 4] - t[0, lo:hi, 3]
print("CTA 0 MMA duration per entry (us):", " ".join(f"{x:.2f}" for x in d))
if os.environ.get("BRIEF"): sys.exit(0)
# cross-CTA: when is entry e's counter complete (max signal over sampled CTAs) vs when dependants see it
DD = int(os.environ.get('DEPDIST', 2))
sig = np.nanmax(t[:, :, 7], axis=0)
print("entry: last sampled signal -> per-CTA depOK of entry+DEPDIST (us)")
for e in range(lo, hi - DD):
    print(f"{E0 + e:5d} {sig[e]:8.2f} -> " + " ".join(f"{t[c, e + DD, 1]:8.2f}" for c in range(NC)))
