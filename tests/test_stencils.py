"""The generated stencil strings are the reference's, token for token (whitespace aside)."""

import json
import os

import pytest

from absinthe_b200 import stencils as S

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "stencil_fingerprints.json")))


def tables():
    out = {"diffusion": S.diffusion_stencils(), "advection": S.advection_stencils(),
           "fastwaves": S.fastwaves_stencils()}
    out.update({v: S.fitcache_stencils(v) for v in S.FITCACHE_VARIANTS})
    out.update({v: S.fitddr_stencils(v) for v in S.FITDDR_VARIANTS})
    return out


@pytest.mark.parametrize("family", sorted(GOLDEN))
def test_strings_match_reference_fingerprints(family):
    mine = tables()[family]
    assert sorted(mine) == sorted(GOLDEN[family])
    for name, code in mine.items():
        assert S.fingerprint(code) == GOLDEN[family][name], (family, name)


def test_sequences_cover_all_stencils():
    assert sorted(S.diffusion_sequence()) == sorted(S.diffusion_stencils())
    assert sorted(S.advection_sequence()) == sorted(S.advection_stencils())
    assert sorted(S.fastwaves_sequence()) == sorted(S.fastwaves_stencils())


def test_fitddr_k88_quirk():
    # ref fitddr.py:102,113,124,133 -- the last k-field is k88, not k8
    assert "k88(i,j,k)" in S.fitddr_stencils("IN3BD0")["o8"]
    assert "k8(" not in S.fitddr_stencils("IN3BD2")["o8"]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources not mounted")
def test_against_live_reference():
    import importlib
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        ref = importlib.import_module("fastwaves")
        norm = lambda s: "".join(s.split())
        for name, code in S.fastwaves_stencils().items():
            assert norm(code) == norm(ref.STENCILS[name])
    finally:
        sys.path.remove("/root/reference")
