"""The reference's own unit tests (35 known-answer vectors) restated against
the CPU oracle at the reference's own board sizes. Bodies live in tests/reference_kat.py."""
import pytest

import reference_kat as K
from engine_facade import OracleEngine


def make(width, height):
    return OracleEngine(width, height)


@pytest.mark.parametrize("name", K.ALL_KATS)
def test_oracle_kat(name):
    getattr(K, name)(make)


@pytest.mark.parametrize("name", K.ALL_KATS)
def test_oracle_kat_on_the_product_board(name):
    """Same bodies on 15x10 (the board the CUDA product fixes) -- proves the restated
    positions are board-size independent before they are used to check the GPU."""
    getattr(K, name)(lambda w, h: OracleEngine(15, 10))
