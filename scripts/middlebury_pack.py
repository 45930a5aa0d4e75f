"""Pack the Middlebury sequences the reference ships (middlebury/<name>/frame10.png, frame11.png, flow10.flo) into
data/_middlebury/<name>.npz so that GPU-box runs (which have no /root/reference) can use them.  The output directory is
git-ignored (it is the reference's data, not ours) but travels with gpurun.  Run in the build container:
    python scripts/middlebury_pack.py [/root/reference/middlebury]
Stored: rgb frames as uint8 (the grey conversion is done by the package's rgb2gray at run time, like optical_flow.m:10-11) and
the raw .flo contents (float32, unknown flow still > 1e9)."""
import os, sys
import numpy as np
from PIL import Image
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/middlebury"
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "_middlebury")
os.makedirs(dst, exist_ok=True)
NAMES = {"RubberWhale": "rubberwhale"}
for name in ("RubberWhale", "Dimetrodon", "Hydrangea", "Venus", "Grove2", "Grove3", "Urban2", "Urban3", "Teddy", "Cones"):
    d = os.path.join(src, NAMES.get(name, name))
    f10 = np.asarray(Image.open(os.path.join(d, "frame10.png")))
    f11 = np.asarray(Image.open(os.path.join(d, "frame11.png")))
    with open(os.path.join(d, "flow10.flo"), "rb") as f:
        tag = np.fromfile(f, np.float32, 1)[0]
        w, h = np.fromfile(f, np.int32, 2)
        flo = np.fromfile(f, np.float32).reshape(h, w, 2)
    assert tag == np.float32(202021.25) and f10.shape[:2] == (h, w)
    np.savez_compressed(os.path.join(dst, name + ".npz"), frame10=f10, frame11=f11, flow10=flo)
    print(name, f10.shape, f10.dtype, flo.shape, "%.1f MB" % (os.path.getsize(os.path.join(dst, name + ".npz")) / 1e6))
