"""TEST / BENCHMARK INFRASTRUCTURE ONLY.  Stages the files of the UNMODIFIED Python reference that the hot
path consists of under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun), so that
`bench.py --impl reference` and its `cpu_baseline` leg can time the reference's own numpy implementation
on the benchmark box, where /root/reference does not exist.  The reference has no packaging files, so the
`pip install --target baseline/_ref` route of the bench contract does not apply: the files are copied
byte for byte (checked by sha256 in the manifest this script writes).

  python oracle/stage_reference.py [reference_root]
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "envs/gym_swimmer/swimmer/remy_swimmer_env.py", "envs/gym_swimmer/swimmer/__init__.py",
    "ars/parameters.py", "ars/environment.py", "ars/ars_agent.py", "ars/database.py", "ars/estimator.py",
    "safe_ars/ars.py", "rlglue/agent/SwimmerAgent.py",
]


def stage(ref="/root/reference"):
    if not os.path.isfile(os.path.join(ref, FILES[0])):
        print("reference tree absent at %s: keeping the staged copy (if any)" % ref)
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(ref, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": ref, "sha256": manifest}, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    print("staged %d reference files under %s" % (len(FILES), DEST))
    return True


if __name__ == "__main__":
    stage(*sys.argv[1:2])
