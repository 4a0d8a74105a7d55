"""ORACLE — test infrastructure only.

CPU restatement (torch fp64/fp32) of the reference hot path plus a numpy stand-in for the
TensorFlow 1.3 primitives (tf_shim) that lets the reference's own Python source run here.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (vmp_for_svae_b200) never does.
"""
