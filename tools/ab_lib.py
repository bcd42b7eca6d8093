"""Run bench.py against another build of the library (same-box A/B of a kernel change):

    python tools/ab_lib.py vsr_b200/lib/libvsr_sm100_head.so --steps 20 --warmup 5 --no-cpu-baseline --no-extras
"""
import os
import runpy
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import vsr_b200.build as b  # noqa: E402

b.LIBPATH = os.path.abspath(sys.argv[1])
b.up_to_date = lambda: True
sys.argv = [os.path.join(root, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
