"""A/B of library builds (BSG_B200_LIB) over every env type: scripts/kernel_time.py once per library."""
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:]:
    print("==", lib, flush=True)
    subprocess.run([sys.executable, os.path.join(root, "scripts", "kernel_time.py")],
                   env=dict(os.environ, BSG_B200_LIB=os.path.join(root, lib)))
