"""Development aid: build copies of libtheta_b200.so whose n=2 search kernel is compiled with different -D switches.

    python tools/build_variants.py name1:-DTHETA_P1_REFINE=0,-DTHETA_P1_INCREMENTAL=0 name2:...

writes theta_b200/variants/libtheta_<name>.so (git-ignored, shipped to the GPU box); run a tool with
THETA_LIB=theta_b200/variants/libtheta_<name>.so to measure it.
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from theta_b200 import build as B

B.build()
out = os.path.join(B.HERE, "variants")
os.makedirs(out, exist_ok=True)
nvcc = B._nvcc()
procs = []
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    src = "theta_n2.cu"
    if "@" in name:
        name, src = name.split("@")
    obj = os.path.join(out, "%s_%s.o" % (src.replace(".cu", ""), name))
    cmd = [nvcc] + B.NVCC_FLAGS + [d for d in defs.split(",") if d] + ["-c", os.path.join(B.CSRC, src), "-o", obj]
    procs.append((name, src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, src, obj, p in procs:
    log, _ = p.communicate()
    if p.returncode:
        raise SystemExit("nvcc failed for %s:\n%s" % (name, log))
    open(obj + ".log", "w").write(log)
    objs = [obj if s == src else os.path.join(B.HERE, "build", s.replace(".cu", ".o")) for s in B.SOURCES]
    lib = os.path.join(out, "libtheta_%s.so" % name)
    subprocess.run([nvcc, "-shared", "-cudart", "static", "-o", lib] + objs, check=True)
    print(lib)
