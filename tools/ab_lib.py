"""A/B different builds of libtmc: python tools/ab_lib.py <lib.so> <config> (root pair timing + tree)."""
import os, sys, shutil
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1]
import importlib
sys.path.insert(0, root)
from too_many_cells_b200 import _lib
_lib.LIB_PATH = os.path.join(root, "too-many-cells_b200", lib)
sys.argv = [sys.argv[0]] + sys.argv[2:]
exec(open(os.path.join(root, "tools", "prof_pair.py")).read())
