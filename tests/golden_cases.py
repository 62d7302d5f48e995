"""Definitions of the golden-fixture cases, shared by tests/make_golden.py (which asks the REFERENCE for the
answers) and the tests (which ask the oracle / the GPU path).  Inputs are regenerated from seeds by the
deterministic synthetic generator, so only the reference's outputs are committed."""
import numpy as np

from workloads import synth

LOADER_FRAME = 2


def loader_systems():
    return {"water": synth.water_box(300, 2.1, seed=11), "il": synth.ionic_liquid(60, 2.9, seed=12)}


def edge_fixture():
    """4 molecules of 3 atoms in a 4 x 5 x 6 box: straddling each face, and an atom exactly L/2 from its reference
    atom (strict inequalities in gmx_handle.ipp:139-142 leave it unshifted)."""
    L = np.array([4.0, 5.0, 6.0], dtype=np.float32)
    x = np.array([
        [3.9, 2.0, 3.0], [0.1, 2.1, 3.1], [3.8, 1.9, 2.9],      # straddles +x face
        [1.0, 0.05, 3.0], [1.1, 4.95, 3.1], [0.9, 0.15, 2.9],   # straddles -y face
        [1.0, 2.0, 5.9], [1.1, 2.1, 0.2], [0.9, 1.9, 5.7],      # straddles +z face
        [0.5, 1.0, 1.0], [2.5, 3.5, 4.0], [2.5000002, 3.5000002, 4.0000005],   # exactly L/2 away, and just beyond
    ], dtype=np.float32)
    box = np.zeros(9, np.float32)
    box[0], box[4], box[8] = L
    idx = np.arange(12, dtype=np.int32)
    mass = np.array([15.9994, 1.008, 1.008], dtype=np.float32).astype(np.float64)
    charge = np.array([-0.8476, 0.4238, 0.4238], dtype=np.float32).astype(np.float64)
    return x, box, idx, 3, mass, charge


def periodicity_inputs():
    rng = np.random.default_rng(5)
    out = [(2.0, 4.0), (-2.0, 4.0), (2.0000000000000004, 4.0), (6.1, 4.0), (-10.3, 4.0), (0.0, 3.3)]
    for _ in range(200):
        L = float(np.float32(rng.uniform(1, 20)))
        out.append((float(rng.uniform(-3 * L, 3 * L)), L))
    return out


def program_cases():
    return {
        "rdf_water_oo": dict(exe="calc_rdf_region", system=lambda: synth.water_box(1200, 3.3, seed=21, split_atoms=True),
                             nframes=3, groups=["OW", "OW"], args=[]),
        "rdf_il_slabs": dict(exe="calc_rdf_region", system=lambda: synth.ionic_liquid(250, 4.3, seed=22), nframes=3,
                             groups=["CAT", "ANI"],
                             args=["-nbin", 3, "-low", 0.5, "-up", 3.9, "-dim", 1, "-dim2", 0, "-low2", 0.3, "-up2", 4.1,
                                   "-Rm", 1.7]),
        "density_template_c1": dict(exe="readme_template", system=lambda: synth.water_box(3000, 4.476, seed=20200101),
                                    nframes=10, groups=["SOL"], args=["-nbin", 100, "-dim", 2, "-low", 0, "-up", 4.476]),
        # general/calc_density_region.cpp cannot supply a fixture: the reference binary aborts at :76 with
        # fmt::format_error ("{:10d}" applied to a double).  Its binning idiom (:59-71) is the README template's.
        "orient_water": dict(exe="calc_orient_mof", system=lambda: synth.water_box(2500, 4.2, seed=24), nframes=4,
                             groups=["SOL"],
                             args=["-NCM", 0, "-dim", 2, "-low", 0, 0, 0.5, "-up", 100, 100, 3.7, "-c1", 2.1, "-c2", 2.1,
                                   "-dr", 0.25, "-CR", 2.0, "-Cr", 0.0, "-nA", 90, "-tp1", 1, "-tp2", 2, "-tp3", 3,
                                   "-DIMNOrient", 2, "-method", 1]),
        # (CR - Cr) / dr must be integral: otherwise R in the last partial bin gives meZ == nbin1 and the reference
        # itself writes out of bounds (calc_orient_mof.cpp:179-180,196), aliasing into orient(0, meA + 1).
        "orient_il_cross": dict(exe="calc_orient_mof", system=lambda: synth.ionic_liquid(300, 4.0, seed=25), nframes=3,
                                groups=["CAT"],
                                args=["-NCM", 1, "-dim", 0, "-low", 0.2, 0, 0, "-up", 3.8, 100, 100, "-c1", 2.0, "-c2", 2.0,
                                      "-dr", 0.5, "-CR", 2.0, "-Cr", 0.5, "-nA", 45, "-tp1", 3, "-tp2", 7, "-tp3", 12,
                                      "-DIMNOrient", 1, "-method", 2]),
        # coordination numbers (mofs/calc_pair_dist_bulk.cpp): anions within Rmin of every cation of a slab
        "pair_dist_il": dict(exe="calc_pair_dist_bulk", system=lambda: synth.ionic_liquid(300, 4.4, seed=26), nframes=4,
                             groups=["CAT", "ANI"], args=["-ncm", 0, "-Rmin", 0.8, "-dim", 2, "-low", 0.6, "-up", 3.9]),
    }


def field_fixture():
    """Velocities / forces for the loaders that do not unwrap (gmx_handle.ipp:201-339): any float array does."""
    system = synth.ionic_liquid(60, 2.9, seed=12)
    rng = np.random.default_rng(17)
    v = rng.normal(size=(system.natoms, 3)).astype(np.float32)
    f = (rng.normal(size=(system.natoms, 3)) * 300.0).astype(np.float32)
    return system, v, f


def full_fixture():
    """Heterogeneous group for GmxHandleFull: all cations (25 atoms) followed by all anions (7 atoms) of a small IL box."""
    system = synth.ionic_liquid(70, 3.0, seed=13)
    napm = np.array([25] * 70 + [7] * 70, dtype=np.int32)
    m, q, *_ = system.atom_tables()
    return system, system.groups["System"], napm, m.astype(np.float64), q.astype(np.float64)
