"""Write a scene container and a console script for the C++ host binary (raym0nade_b200/raym0nade):

    python scripts/console_demo.py [out_dir]            # BASELINE configs[0]: Cornell box, 512x512, 64 spp
    ./raym0nade_b200/raym0nade < out_dir/script.txt     # on a B200 box; PNGs land in gpurun_out/console/

The script is the reference console's own dialogue (src/myconsole.cpp): create model / create args / render.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raym0nade_b200 import scenes  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "_console_demo"
os.makedirs(out, exist_ok=True)
scene, a = scenes.cornell_box(512, 512, 64)
scene.save(os.path.join(out, "cornell.rmscene"))
g = lambda v: " ".join("%.9g" % np.float32(x) for x in v)
d, r, u, p = (np.float32(a.direction), np.float32(a.right), np.float32(a.up), np.float32(a.position))
args = "\n".join([g(d), g(r), g(u), g([np.dot(p, d), np.dot(p, r), np.dot(p, u)]),      # position as (D, R, U) coefficients
                  g([a.accuracy, 3.4, 6.0, a.exposure]), "512 512", "64 8 0.7", "gpurun_out/console/cornell"])
with open(os.path.join(out, "script.txt"), "w") as f:
    f.write("create model box\n%s/\ncornell.rmscene\nnull\nview model box\ncreate args cfg0\n%s\nview args cfg0\nrender box cfg0\nexit\n" % (out, args))
print(os.path.join(out, "script.txt"))
