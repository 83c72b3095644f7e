"""Debugging aid: does freeing a Bilateral filter return its device LUTs?  Run under compute-sanitizer --tool memcheck --leak-check full."""
import ctypes as C, gc, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import vapoursynth_zip_b200 as vz
from helpers import noise_clip, to_node
vz.core.init([0])
lib = vz.load_library()
clip = noise_clip("YUV420P16", 320, 180, seed=1)
mode = sys.argv[1] if len(sys.argv) > 1 else "del"
node = to_node(clip).vszip.Bilateral(sigmaS=2, sigmaR=2)
node.get_frame(0)
flt = node.filter
if mode == "explicit":
    h, flt.handle = flt.handle, None
    lib.vszip_filter_free(h)
    print("explicit free done, error:", repr(vz._last_error()))
else:
    calls = []
    orig = vz._Filter.__del__
    def dbg(self):
        calls.append((type(self).__name__, self.handle)); orig(self)
    vz._Filter.__del__ = dbg
    del node, flt
    gc.collect()
    print("__del__ calls:", calls, "error:", repr(vz._last_error()))
vz.core.shutdown()
print("done")
