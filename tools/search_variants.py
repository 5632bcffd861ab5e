"""Pass-1 search kernel at each THETA_SEARCH_MINB (read once per process -> subprocesses). Development aid."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in ("6", "5", "4"):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench.py")] + (sys.argv[1:] or ["cfg2", "m150k7"]),
                   env=dict(os.environ, THETA_SEARCH_MINB=v))
