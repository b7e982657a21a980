"""The generated product rows of fe_mul (csrc/fe_mul_rows.inc, the -DS252_FE_MUL_GEN=1 variant): the instruction table simulates
correctly against python integers (no lost carry on edge operands) and the committed file is what the generator emits."""
import importlib.util
import os

from conftest import ROOT


def _gen():
    spec = importlib.util.spec_from_file_location("gen_fe_mul", os.path.join(ROOT, "tools", "gen_fe_mul.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_rows_simulate_correctly():
    g = _gen()
    cases, n_mad = g.selftest()
    assert cases > 2000 and n_mad == 64


def test_committed_rows_are_current():
    g = _gen()
    path = os.path.join(ROOT, "lambdaworks_cairo_prover_b200", "csrc", "fe_mul_rows.inc")
    assert open(path).read().strip() == g.emit().strip()
