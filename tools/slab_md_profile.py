import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from apax_b200 import synthetic as syn
from apax_b200.md.ase_calc import Atoms
from apax_b200.md import domain_md
from apax_b200.md.domain_md import SlabNVT
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pos, Z, cell = syn.water_box(333334, 0)
n = len(Z); masses = np.where(Z == 8, 15.999, 1.008)
fs = 0.09822694788464063; dt, kT = 0.25 * fs, 300.0 * 8.617333262e-5
p0 = np.random.default_rng(0).normal(size=(n, 3)) * np.sqrt(masses * kT)[:, None]; p0 -= p0.mean(0)
md = SlabNVT(syn.gmnn_params(seed=1234), Atoms(positions=pos, numbers=Z, cell=cell), None, dt=dt, kT=kT, dr_threshold=0.5, rebuild_every=10, masses=masses, momentum=p0)
T = {"gather": 0.0, "redistribute": 0.0}
og, orr = md._gather, md._redistribute
def g():
    torch.cuda.synchronize(); t = time.perf_counter(); og(); torch.cuda.synchronize(); T["gather"] += time.perf_counter() - t
def r(first=False):
    torch.cuda.synchronize(); t = time.perf_counter(); orr(first); torch.cuda.synchronize(); T["redistribute"] += time.perf_counter() - t
md._gather, md._redistribute = g, r
md.profile = {}
# FINE: time the collective inside configure separately
import torch.distributed as _d
_orig_ag = _d.all_gather_into_tensor
def _ag(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = _orig_ag(*a, **k); torch.cuda.synchronize()
    md.profile["cfg_allgather"] = md.profile.get("cfg_allgather", 0.0) + time.perf_counter() - t0; return r
_d.all_gather_into_tensor = _ag
md.run(10); torch.cuda.synchronize(); dist.barrier()
T = {k: 0.0 for k in T}; md.profile.clear()
t0 = time.perf_counter(); md.run(30); torch.cuda.synchronize(); dist.barrier(); wall = time.perf_counter() - t0
print(json.dumps({"rank": rank, "wall_ms_per_step": wall / 30 * 1e3, **{k + "_ms_per_rebuild": v / 3 * 1e3 for k, v in T.items()}, **{"redis_" + k: v / 3 * 1e3 for k, v in md.profile.items()}}))
# un-instrumented: cost of one re-decomposition = (wall with a rebuild every 10 steps - wall with none) * 10
md.profile = None; md._gather, md._redistribute = og, orr; _d.all_gather_into_tensor = _orig_ag
walls = {}
for every in (10, 1000000):
    md.rebuild_every = every
    md.run(10); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); md.run(60); torch.cuda.synchronize(); dist.barrier(); walls[every] = (time.perf_counter() - t0) / 60 * 1e3
if rank == 0:
    print(json.dumps({"uninstrumented_ms_per_step_rebuild10": walls[10], "no_rebuild": walls[1000000],
                      "ms_per_rebuild": (walls[10] - walls[1000000]) * 10}))
md.rebuild_every = 10
if os.environ.get("SLAB_CPROFILE"):
    import cProfile, pstats, io
    md.profile = None; md._gather, md._redistribute = og, orr; _d.all_gather_into_tensor = _orig_ag
    pr = cProfile.Profile(); pr.enable(); md.run(30); torch.cuda.synchronize(); pr.disable()
    if rank == 0:
        st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28); print(st.getvalue())
md.close(); dist.destroy_process_group()
