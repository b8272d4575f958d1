import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import euclider_b200 as eb
name = sys.argv[1] if len(sys.argv) > 1 else "3d_fresnel"
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 48)
env = eb.load_reference_scene(name)
if len(sys.argv) > 4 and sys.argv[4] == "mega":
    env.pipeline = eb.EUCL_PIPELINE_MEGAKERNEL
img = env.render((w, h), time=0.0, want_hit_ids=True)
print(img.stats)
