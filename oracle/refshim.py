"""Import the LIVE reference (/root/reference/python) under Python 3.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) validate the restatement
in oracle/theta_oracle.py against the reference's own code and (b) generate the
golden vectors committed under tests/golden/ (see oracle/make_golden.py), and
(c) time the reference's own multiprocessing path as the CPU arm of bench.py.
/root/reference exists only in the build container; on the GPU box the modules
come from oracle/_ref/python, a git-ignored snapshot made by oracle/build_ref.sh
that travels with the working tree.  Only bench.py's CPU arm uses it there.

No reference source is copied: each module's text is read from the read-only
tree at import time and executed after the minimal py2->py3 syntax rewrites
listed in SURVEY.md Appendix C (print statements, string.join, time.clock,
one tuple-parameter lambda).  No arithmetic line is touched.
"""
import importlib.abc
import importlib.util
import os
import re
import sys
import types

_SNAP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "python")  # oracle/build_ref.sh (git-ignored)
REF = os.environ.get("THETA_REFERENCE") or ("/root/reference/python" if os.path.isdir("/root/reference/python") else _SNAP)

_PRINT = re.compile(r"^(\s*(?:[^#'\"]*?:\s*)?)print\s+(?!\()(.*)$")
_BARE_PRINT = re.compile(r"^(\s*(?:[^#'\"]*?:\s*)?)print\s*$")


def _rewrite(src):
    out = []
    for line in src.split("\n"):
        s = line.lstrip()
        if not s.startswith("#"):
            m = _PRINT.match(line)
            if m and (m.group(1).strip() == "" or m.group(1).rstrip().endswith(":")):
                body = m.group(2).rstrip()
                if body.endswith(","):
                    line = "%sprint(%s end=' ')" % (m.group(1), body)
                else:
                    line = "%sprint(%s)" % (m.group(1), body)
            else:
                m = _BARE_PRINT.match(line)
                if m and (m.group(1).strip() == "" or m.group(1).rstrip().endswith(":")):
                    line = m.group(1) + "print()"
        line = re.sub(r'string\.join\((.*),\s*","\)', r'",".join(\1)', line)
        line = line.replace("time.clock()", "time.process_time()")
        line = line.replace("lambda (num, denom): num / denom if denom > 0 else -1",
                            "lambda nd: nd[0] / nd[1] if nd[1] > 0 else -1")
        out.append(line)
    return "\n".join(out)


_STUBS = {
    "plotResults": ["plot_results"],
    "RunBAFModel": ["run_BAF_model", "calculate_BAF", "generate_pi", "calculate_interval"],
    "ClusteringBAF": ["clustering_BAF", "group_to_meta_interval"],
}
_REWRITTEN = {"RunTHetA", "DataTools", "TimeEstimate", "SelectIntervals", "FileIO",
              "ModelSelection", "CalcAllC"}


class _Loader(importlib.abc.Loader):
    def __init__(self, name):
        self.name = name

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        path = os.path.join(REF, self.name + ".py")
        with open(path) as f:
            src = _rewrite(f.read())
        module.__file__ = path
        exec(compile(src, path, "exec"), module.__dict__)


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name in _REWRITTEN and os.path.exists(os.path.join(REF, name + ".py")):
            return importlib.util.spec_from_loader(name, _Loader(name))
        return None


_installed = False


def install():
    """Make `import RunTHetA, Optimizer, Enumerator, CalcAllC, ...` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present at %s (build container only)" % REF)
    sys.dont_write_bytecode = True
    for mod, names in _STUBS.items():
        m = types.ModuleType(mod)
        for n in names:
            def _raise(*a, _n=n, **k):
                raise RuntimeError("stubbed reference function %s called" % _n)
            setattr(m, n, _raise)
        if mod == "plotResults":
            m.plot_results = lambda *a, **k: None
        sys.modules[mod] = m
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REF)
    _installed = True


def modules():
    install()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import Misc, Enumerator, Optimizer, CalcAllC, DataTools, TimeEstimate, SelectIntervals, FileIO, RunTHetA
        import ModelSelection
    return types.SimpleNamespace(Misc=Misc, Enumerator=Enumerator, Optimizer=Optimizer,
                                 CalcAllC=CalcAllC, DataTools=DataTools, TimeEstimate=TimeEstimate,
                                 SelectIntervals=SelectIntervals, FileIO=FileIO, RunTHetA=RunTHetA,
                                 ModelSelection=ModelSelection)
