"""n=3 search kernel at each THETA_N3_BLOCK (read once per process -> subprocesses). Development aid."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in (sys.argv[1:] or ["64", "256", "512"]):
    print("THETA_N3_BLOCK", v, flush=True)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_n3.py")], env=dict(os.environ, THETA_N3_BLOCK=v))
