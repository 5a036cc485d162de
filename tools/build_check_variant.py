"""Build the checking variant of the library used by tools/gpu_sanitize.sh (racecheck): every lane arrives on the mbarriers."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lpcnet_b200 import build
print(build.build_variant("arriveall", ["LPCNET_ARRIVE_ALL=1"]))
