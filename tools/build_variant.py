"""Kernel experiments: build the plugin with extra -D flags into acts_b200/variants/<name>.so.

    python tools/build_variant.py gap0 -DB200SEED_GAP_LOOKBACK=0
    B200SEED_LIB=acts_b200/variants/gap0.so python profiles/profile_driver.py --events 16
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acts_b200 import build  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(build.ROOT, "acts_b200", "variants", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
build.OBJ_DIR = os.path.join(build.ROOT, "build", "obj_" + name)
cmds = build.engine_compile_commands(extra=flags)
procs = [subprocess.Popen(cmd) for _, cmd in cmds]
if any(p.wait() != 0 for p in procs):
    sys.exit("compile failed")
subprocess.run(["nvcc", *build.COMMON_FLAGS, "-shared", "-o", out, *[obj for obj, _ in cmds]], check=True)
print(out)
