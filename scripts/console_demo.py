"""Write a scene container and a console script for the C++ host binary (raym0nade_b200/raym0nade):

    python scripts/console_demo.py [out_dir] [--config 0|1|2] [--spp N]
    ./raym0nade_b200/raym0nade < out_dir/script.txt     # on a B200 box; PNGs land in gpurun_out/console/

  --config 0   BASELINE configs[0]: Cornell box, 512x512, 64 spp (default)
  --config 1   configs[1]: 260 K triangles + HDR sky (embedded in the .rmscene), 1080p, 256 spp
  --config 2   configs[2]: the 1 M-triangle glossy / dielectric scene, 1080p, 1024 spp (the bench workload; ~100 MB file)

The script is the reference console's own dialogue (src/myconsole.cpp): create model / create args / render.  The camera
position is given as (D, R, U) coefficients of direction / right / up, as the console asks for it (src/myconsole.cpp:47-51).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raym0nade_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("out", nargs="?", default="_console_demo")
ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2])
ap.add_argument("--spp", type=int, default=None)
opt = ap.parse_args()
os.makedirs(opt.out, exist_ok=True)
if opt.config == 0:
    name, (scene, a), lens = "cornell", scenes.cornell_box(512, 512, 64), (3.4, 6.0)      # focus / CoC: the depth-of-field exports too
elif opt.config == 1:
    name, (scene, a), lens = "sponza", scenes.sponza_scale(), (0.0, 0.0)
else:
    name, (scene, a), lens = "glossy", scenes.glossy_dielectric(), (0.0, 0.0)
spp = a.spp if opt.spp is None else opt.spp
scene.save(os.path.join(opt.out, name + ".rmscene"))
g = lambda v: " ".join("%.9g" % np.float32(x) for x in v)
d, r, u, p = (np.float32(a.direction), np.float32(a.right), np.float32(a.up), np.float32(a.position))
args = "\n".join([g(d), g(r), g(u), g([np.dot(p, d), np.dot(p, r), np.dot(p, u)]),
                  g([a.accuracy, lens[0], lens[1], a.exposure]), "%d %d" % (a.width, a.height), "%d 8 %s" % (spp, g([a.P_Direct])),
                  "gpurun_out/console/" + name])
with open(os.path.join(opt.out, "script.txt"), "w") as f:
    f.write("create model m\n%s/\n%s.rmscene\n%s\nview model m\ncreate args cfg\n%s\nview args cfg\nrender m cfg\nexit\n"
            % (opt.out, name, "embedded" if scene.sky is not None else "null", args))
print(os.path.join(opt.out, "script.txt"))
