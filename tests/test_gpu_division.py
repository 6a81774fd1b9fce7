"""csrc/fastdiv.cuh (shared-reciprocal division used by the pivot kernels) must have the bits of IEEE division
(__ddiv_rn == JS `/`, src/simplex.ts:19,25,36,89,128) on every operand pair, special values included."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_fastdiv_matches_ddiv_rn(engine, mode):
    n = 1 << 29 if mode in (0, 1) else 1 << 27
    for seed in (1, 0xDEADBEEF):
        bad, first = engine.probe_division(n, seed, mode)
        assert bad == 0, f"mode {mode}: {bad} mismatches, first n={first[0]:#018x} d={first[1]:#018x}"
