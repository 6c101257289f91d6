"""Kernel experiments: build the plugin with extra -D flags into acts_b200/variants/<name>.so.

    python tools/build_variant.py gap0 -DB200SEED_GAP_LOOKBACK=0
    B200SEED_LIB=acts_b200/variants/gap0.so python profiles/profile_driver.py --events 16
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acts_b200.build import CSRC, NVCC_FLAGS, ROOT  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "acts_b200", "variants", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = ["nvcc", *NVCC_FLAGS, *flags, "-Xptxas", "-warn-spills", "-o", out,
       os.path.join(CSRC, "seeding_plugin.cu"), os.path.join(CSRC, "host_plan.cpp")]
subprocess.run(cmd, check=True)
print(out)
