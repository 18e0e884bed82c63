"""The register selection code of the coordinate-wise CUDA kernels (csrc/cw_core.cuh: odd-even merge network,
median-pinning padding, trimmed mean, mean-of-medians, synthesised rows, NaN canonicalisation, and the PREPAD
form used by the warp-tiled kernel) is ``__host__ __device__``: this test compiles it for the HOST with nvcc
and runs ~1M randomised / exhaustive cases against ``std::sort`` -- the same source the kernels compile, checked
on a box without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("cw_host") / "cw_network_host")
    res = subprocess.run([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "byzpy_b200", "csrc"),
                          os.path.join(ROOT, "tests", "native", "cw_network_host.cu"), "-o", exe],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    return exe


def test_selection_code_matches_a_sort_for_every_size_mode_and_padding(harness):
    res = subprocess.run([harness, "3"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:]
    last = res.stdout.strip().splitlines()[-1]
    assert last.endswith("0 failures") and int(last.split()[0]) > 500_000, last


def _tiled_addresses(NP, n, lanes=32):
    """Python model of cw_select_tiled_kernel's index maps (csrc/cw_select.cu): which (row, coordinate)
    every lane's 16-byte cp.async covers, and where it lands in the warp tile."""
    writes = {}
    for j in range(NP // 4):
        for lane in range(lanes):
            sub, chunk = lane >> 3, (lane & 7) * 4
            r = j * 4 + sub
            if r < n:
                for c in range(4):
                    writes[(r * 32 + chunk + c)] = (r, chunk + c)         # tile offset -> (row, coordinate)
    return writes


@pytest.mark.parametrize("NP,n", [(32, 17), (32, 32), (64, 33), (64, 50), (64, 64), (128, 65), (128, 128)])
def test_tiled_kernel_index_maps_cover_every_real_row_exactly_once_without_bank_conflicts(NP, n):
    w = _tiled_addresses(NP, n)
    # every (row < n, coordinate < 32) is written exactly once, at tile offset row * 32 + coordinate
    assert len(w) == n * 32 and all(off == r * 32 + c for off, (r, c) in w.items())
    # read side: lane l reads offset i * 32 + l for i = 0..NP-1 -> bank (i * 32 + l) % 32 = l: conflict free
    for i in range(NP):
        assert sorted((i * 32 + lane) % 32 for lane in range(32)) == list(range(32))
    # write side: a 16-byte cp.async is served per quarter-warp (8 lanes): 8 chunks of one row = 32 distinct banks
    for lane0 in range(0, 32, 8):
        banks = [((lane0 >> 3) * 32 + (lane & 7) * 4 + c) % 32 for lane in range(lane0, lane0 + 8) for c in range(4)]
        assert sorted(banks) == list(range(32))
