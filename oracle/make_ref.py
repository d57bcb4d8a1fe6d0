#!/usr/bin/env python
"""Places the reference's own Python modules for the hot path under oracle/_ref/ (TEST INFRASTRUCTURE).

The reference is pure Python, so "building" it is a copy of the few files the path lives in, from
where they lie in the reference checkout.  oracle/_ref/ is git-ignored (no reference source enters
the history) but travels to the GPU box with the snapshot, where /root/reference does not exist:
  * bench.py --impl reference times gnn/model.py's SegmentClassifier itself (cpu_baseline.kind
    "reference") instead of the oracle's restatement,
  * tests/test_gpu_store.py drives the drop-in module with the reference's own Estimator.
Nothing under gnn_fpga_b200/ ever imports it.  Without the checkout this script does nothing.
"""
import os
import shutil
import sys

REF = os.environ.get("GNNSEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["gnn/model.py", "gnn/model_maskedlinear.py", "gnn/estimator.py"]


def main():
    if not os.path.isdir(os.path.join(REF, "gnn")):
        print("make_ref: no reference checkout at %s; oracle/_ref left as it is" % REF)
        return 0
    out = os.path.join(HERE, "_ref")
    os.makedirs(out, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(out, os.path.basename(f)))
    with open(os.path.join(out, "README"), "w") as fh:
        fh.write("Unmodified copies of %s from the reference checkout (oracle/make_ref.py). Not tracked.\n" % ", ".join(FILES))
    print("make_ref: placed %d reference modules under %s" % (len(FILES), out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
