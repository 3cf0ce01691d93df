"""Host-side cost of one fused render step (python + ctypes + launches), GPU box only.
usage: python tools/host_profile.py"""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
from types import SimpleNamespace
from adgs_b200.gaussian_renderer import render

dev = torch.device("cuda", 0)
wl = B.WORKLOADS["kitti-375x1242-1M"]
model, n_scene, n_obj = B.build_ours(wl, dev)
cam, t, flow_t = B.view_for_rank(wl, 0, dev)
cot = B.make_cotangents(wl, dev)
pipe = SimpleNamespace(inv_depth=True, debug=False, sync_free=True)
flow_pkg = [flow_t, None, None, None, None, None]
params = model.hot_parameters()

def step():
    for p in params:
        p.grad = None
    res = render(cam, model, None, pipe, flow_pkg=flow_pkg, render_objmask=True)
    outs = (res["render"], res["depth"], res["img_opacity"], res["img_flow"], res["img_semantic"])
    cots = (cot["color"], cot["depth"][0], cot["opacity"][0], cot["flow"], cot["semantic"])
    torch.autograd.backward(outs, cots)

for _ in range(5):
    step()
torch.cuda.synchronize()
# pure host time per step when the GPU is idle at the start (sync each step, like e2e)
ts = []
for _ in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
ts.sort()
print("host launch time / total (median): %.3f ms / %.3f ms" % (ts[15][0] * 1e3, sorted(x[1] for x in ts)[15] * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue()[:6000])
