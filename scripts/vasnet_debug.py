"""Runs one VASNet case per subprocess with SMZ_DEBUG_SYNC=1 to localise a failing step."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle.gen_golden_models import build_vasnet, make_input
from summarizer_b200.models.vasnet import VASNet
T, B = int(sys.argv[1]), int(sys.argv[2])
m = build_vasnet(VASNet, 0, {}, 1.0).cuda()
x = make_input(0, T, B).cuda()
with torch.no_grad():
    y = m(x)
torch.cuda.synchronize()
print("ok", T, B, float(y.mean()))
''' % ROOT
for T, B in [(10, 3), (10, 1), (64, 1), (30, 1), (128, 1), (130, 1), (8, 1), (16, 2), (300, 1)]:
    env = dict(os.environ, SMZ_DEBUG_SYNC="1")
    r = subprocess.run([sys.executable, "-c", CHILD, str(T), str(B)], env=env, capture_output=True, text=True, timeout=120)
    tail = (r.stdout.strip().splitlines() or [""])[-1] + " | " + (r.stderr.strip().splitlines() or [""])[-1]
    print(f"T={T} B={B} rc={r.returncode}: {tail}", flush=True)
