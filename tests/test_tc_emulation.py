"""CPU check of the chunked tensor-core algebra (tests/tc_emulation.py): the same identities the sm_100a
kernels use, with bf16 rounding where they round, against the fp64 oracle -- and the two numerical design
decisions for gw (integer log2 references + bit-identical pair products, direct expansion) shown to matter."""
import torch

import tests.tc_emulation as E


def _gw(decay, shape, seed, **flags):
    old = (E.INT_RHO, E.AB_ROUNDED)
    try:
        E.INT_RHO, E.AB_ROUNDED = flags.pop("int_rho", True), flags.pop("ab_rounded", True)
        return E.report(*shape, decay, seed=seed, **flags)
    finally:
        E.INT_RHO, E.AB_ROUNDED = old


def test_chunked_identities_without_rounding():
    res = _gw("model", (1, 130, 2), 3, bf=False, gl_mode="direct")
    for key, (rel, frac) in res.items():
        assert rel < 2.5e-3 and frac < 0.6, (key, rel, frac)      # only the final bf16 rounding of the outputs


def test_kernel_numerics_are_within_tolerance():
    for decay, shape in (("model", (2, 256, 2)), ("randn", (2, 64, 2))):
        res = _gw(decay, shape, 5, bf=True, gl_mode="direct")
        for key, (rel, frac) in res.items():
            assert rel < 5e-3 and frac < 0.8, (decay, key, rel, frac)


def test_why_integer_references_and_direct_expansion():
    # real-valued references: the pair terms of gw no longer telescope exactly -> strong decays break gw
    naive = _gw("randn", (2, 64, 2), 5, bf=True, gl_mode="ab", int_rho=False, ab_rounded=False)
    good = _gw("randn", (2, 64, 2), 5, bf=True, gl_mode="direct")
    assert naive["gw"][0] > 5 * good["gw"][0]
    # bf16 staging of the intermediate sums: fails the max-abs bound on model-like decays
    staged = _gw("model", (2, 256, 2), 5, bf=True, stage_bf16=True, gl_mode="ab", int_rho=False, ab_rounded=False)
    assert staged["gw"][1] > 1.0 > good["gw"][1]
