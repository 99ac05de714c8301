"""One modal batch (uniform_frame(64), 1024 models, all modes + shapes) for kernel-level timing under ncu."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(json.dumps(bench.modal_leg(0, steps=int(os.environ.get("PROBE_STEPS", "1")))))
