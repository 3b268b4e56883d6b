"""Wide-vs-narrow mapping on the cube-contact scenario: one sub-step from identical states, report differing envs."""
import os
import sys
import ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
variant = os.environ.get("HRLW_VARIANT", "")
os.environ["HRL_B200_LIB"] = os.path.join(ROOT, "hrl_pybullet_envs_b200", "libhrl_b200_dbg%s.so" % variant)
from hrl_pybullet_envs_b200 import _cabi
defs = {"HRL_DEBUG_CONTACTS": 1}
if variant:
    defs["HRLW_" + variant] = 1
_cabi.build(defines=defs, out=os.environ["HRL_B200_LIB"])
import numpy as np
import torch
from hrl_pybullet_envs_b200 import VecEnv, config as K

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 256
kw = dict(robot_coll_dist=0.0)
a = VecEnv("AntGatherBulletEnv-v0", N, seed=4, **kw); b = VecEnv("AntGatherBulletEnv-v0", N, seed=4, **kw)
b.L.hrl_set_lanes_per_env(b.h, lanes)
a.reset(); b.reset()
gen = torch.Generator().manual_seed(9)
for t in range(30):
    a.step((torch.rand(N, 8, generator=gen) * 2 - 1).cuda())
f, i = a.get_state(); f = f.cpu().numpy()
rng = np.random.default_rng(2)
ang = rng.uniform(0, 2 * np.pi, (N, 16)); rad = rng.uniform(0.45, 1.05, (N, 16))
f[:, K.SF_ITEMS:K.SF_ITEMS + 32:2] = f[:, [K.SF_POS]] + rad * np.cos(ang)
f[:, K.SF_ITEMS + 1:K.SF_ITEMS + 32:2] = f[:, [K.SF_POS + 1]] + rad * np.sin(ang)
import itertools
f0 = f.copy()
rep = len(sys.argv) > 2
for nsub, n_near in itertools.product((1,), (0, 2, 4)):
    f = f0.copy()
    if rep:   # every warp of the wide mapping holds 4 (8 lanes) copies of ONE env: no two different envs share a warp
        f = np.repeat(f[:N // 4], 4, axis=0)
    if n_near < 16:   # keep only the first n_near cubes near the ant, park the others far away
        f[:, K.SF_ITEMS + 2 * n_near:K.SF_ITEMS + 32] = 6.5
    print("cubes near:", n_near)
    a.set_state(torch.tensor(f), i); b.set_state(torch.tensor(f), i)
    act = (torch.rand(N, 8, generator=gen) * 2 - 1).cuda()
    dbg = torch.zeros(N, 4, 40, device="cuda")
    a.L.hrl_debug_contacts.argtypes = [C.c_void_p]
    a.L.hrl_debug_contacts(C.c_void_p(dbg.data_ptr())); a.substeps(act, nsub); torch.cuda.synchronize(); da = dbg.clone().cpu().numpy()
    dbg.zero_(); b.substeps(act, nsub); torch.cuda.synchronize(); db = dbg.clone().cpu().numpy()
    a.L.hrl_debug_contacts(None)
    cd = np.abs(da - db).max(axis=(1, 2))
    print("  candidate lists differ in", int((cd > 1e-6).sum()), "envs")
    for e in np.nonzero(cd > 1e-6)[0][:3]:
        for k in range(4):
            if np.abs(da[e, k] - db[e, k]).max() > 1e-6:
                print("   env", e, "leg", k, "narrow n", da[e, k, 0], "wide n", db[e, k, 0])
                for c in range(4):
                    print("      c", c, "narrow", np.round(da[e, k, 1 + 8 * c:9 + 8 * c], 4).tolist(), "| wide", np.round(db[e, k, 1 + 8 * c:9 + 8 * c], 4).tolist())
    sa = a.stats(); sb = b.stats()
    fa, _ = a.get_state(); fb, _ = b.get_state()
    d = (fa - fb).abs()[:, :29].max(dim=1).values.cpu().numpy()
    bad = np.nonzero(d > 1e-4)[0]
    print("nsub", nsub, "stats narrow", sa, "wide", sb)
    print("  differing envs:", len(bad), "of", N, "max diff", d.max())
    percl = da[:, :, 0].astype(int)
    isbad = d > 1e-4
    import collections
    cb = collections.Counter(tuple(r) for r in percl[isbad]); cg = collections.Counter(tuple(r) for r in percl[~isbad])
    print("   per-leg contact counts of differing envs:", cb.most_common(12))
    print("   per-leg contact counts of agreeing envs:", cg.most_common(12))
    cube = (da[:, :, 8::8][:, :, :4] >= 4).sum(axis=(1, 2))
    print("   cube contacts in differing envs:", collections.Counter(cube[isbad].tolist()), "agreeing:", collections.Counter(cube[~isbad].tolist()))
    for e in bad[:5]:
        dd = (fa[e] - fb[e]).abs().cpu().numpy()[:29]
        print("   env", e, "diff idx", np.nonzero(dd > 1e-4)[0].tolist(), dd.max())
