"""theta_b200 -- B200-native THetA likelihood search (drop-in for RunTHetA's hot path).

Public surface mirrors the reference modules:
    theta_b200.search.do_optimization / do_optimization_single   (python/RunTHetA.py:124-220)
    theta_b200.enumerator.Enumerator                             (python/Enumerator.py)
    theta_b200.optimizer.Optimizer                               (python/Optimizer.py)
    theta_b200.calc_all_c.L2 / L3                                (python/CalcAllC.py:44-76)
The compute lives in libtheta_b200.so (include/theta_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
