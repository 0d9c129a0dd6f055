import sys, numpy as np
sys.path.insert(0, '/root/repo')
from raym0nade_b200 import scenes
which = sys.argv[1] if len(sys.argv) > 1 else 'glossy'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
if which == 'glossy': sc, a = scenes.glossy_dielectric(n, 64, 36, 0)
elif which == 'sponza': sc, a = scenes.sponza_scale(n, 64, 36, 0, tex_size=64)
pos = np.ascontiguousarray(sc.positions, np.float32).reshape(-1, 9)
pos.tofile('/tmp/ploc/tris.bin')
rng = np.random.default_rng(1)
m = 200000
t = rng.integers(0, pos.shape[0], m)
P = pos[t].reshape(m, 3, 3)
w = rng.random((m, 3)); w /= w.sum(1, keepdims=True)
org = (P * w[:, :, None]).sum(1)
nrm = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]); nrm /= (np.linalg.norm(nrm, axis=1, keepdims=True) + 1e-20)
d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
flip = (d * nrm).sum(1) < 0; d[flip] *= -1
org = org + 1e-3 * nrm * np.where(((nrm*d).sum(1) > 0), 1, -1)[:, None]
np.concatenate([org, d], 1).astype(np.float32).tofile('/tmp/ploc/rays.bin')
print(pos.shape, m)
