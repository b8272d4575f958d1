import sys, numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import euclider_b200 as eb, oracle_api
env = eb.load_reference_scene(sys.argv[1] if len(sys.argv) > 1 else "3d_fresnel")
env.precision = sys.argv[2] if len(sys.argv) > 2 else "f32"
w, h = 64, 36
import os
if os.environ.get("PROBE_DEPTH"): env.camera.max_depth = int(os.environ["PROBE_DEPTH"])
ref = oracle_api.render(env, w, h, time=0.5, variant=("f32" if env.precision == "f32" else "det"))
img = env.render((w, h), time=0.5, want_hit_ids=True)
print("hit equal", np.array_equal(img.hit_ids, ref[1]), "rgb equal", np.array_equal(img.data, ref[0]), "levels", img.stats["level_counts"] == ref[2]["level_counts"])
print(img.stats["level_counts"], ref[2]["level_counts"])
d = np.abs(img.data.astype(int) - ref[0].astype(int)).max(axis=-1)
print("pixels differing", int((d > 0).sum()), "max diff", int(d.max()), "hit diff", int((img.hit_ids != ref[1]).sum()))
print("gpu", img.data[18, 30:34].tolist(), "ref", ref[0][18, 30:34].tolist())
print("gpu row0", img.data[0, :4].tolist(), "ref", ref[0][0, :4].tolist())
# where do the gpu colours come from?  look for each gpu pixel value in the reference frame
flat_ref = ref[0].reshape(-1, 3)
for idx in (0, 1, 2, 3, 64, 65, 1000):
    px = img.data.reshape(-1, 3)[idx]
    where = np.where((flat_ref == px).all(axis=1))[0][:5]
    print("gpu pixel", idx, px.tolist(), "found in ref at", where.tolist())
