"""Randomised stress of the device paths, a few dozen cases each (the scripts take --cases / --seed for longer runs; results of
the long runs are in profiles/).  Loaders against the oracle; K3 five-way self-consistency (FP32 bracket, every pair in FP64,
cells, symmetric, cells + symmetric); K2 / K4 FP32 brackets in the three modes (literal, bracket + queue, validation)."""
import pytest

import stress_fast_paths
import stress_loaders
import stress_rdf

pytestmark = pytest.mark.gpu


def test_loaders_random_cases_bit_exact_against_oracle():
    assert stress_loaders.main(["--cases", "40", "--seed", "101"]) == 0


def test_rdf_random_cases_five_way_consistency():
    assert stress_rdf.main(["--cases", "30", "--seed", "102"]) == 0


def test_fast_path_brackets_random_cases():
    assert stress_fast_paths.main(["--cases", "40", "--seed", "103"]) == 0
