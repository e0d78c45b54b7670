"""A/B of an alternative build of the engine library (argv[1] = path of the .so) -- same arguments as tools/time_impl.py after it."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sugarscape_utilicalc_b200 import capi  # noqa: E402
capi.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [sys.argv[0]] + sys.argv[2:]
exec(open(os.path.join(ROOT, "tools", "time_impl.py")).read())
