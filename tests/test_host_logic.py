"""Host-side logic shared between the planner and the kernels, compiled with the host compiler only (no GPU, no nvcc):
gemm_warp_grid (quantit_b200/csrc/qtb_core.h) decides how the four consumer warps of a 64 x 64 tile share its valid 8x8
atoms; the kernel and the planner's cost model both call it, so it is checked exhaustively here: every valid atom is
owned by exactly one warp, no warp owns more than 4 x 4 atoms, and no other admissible grid shape puts fewer atoms on the
busiest warp."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include "qtb_core.h"
int main()
{
	int bad = 0;
	for (int mv = 1; mv <= 8; ++mv)
		for (int nv = 1; nv <= 8; ++nv)
		{
			int gm, gn, am, an;
			qtb::gemm_warp_grid(mv, nv, gm, gn, am, an);
			if (gm * gn != 4 || am < 1 || an < 1 || am > 4 || an > 4)
				++bad;
			int owner[8][8] = {};
			for (int w = 0; w < 4; ++w)
			{
				const int wi = w / gn, wj = w - wi * gn;
				int mi = mv - wi * am; mi = mi < 0 ? 0 : (mi > am ? am : mi);
				int nj = nv - wj * an; nj = nj < 0 ? 0 : (nj > an ? an : nj);
				for (int i = 0; i < mi; ++i)
					for (int j = 0; j < nj; ++j)
						owner[wi * am + i][wj * an + j] += 1;
			}
			for (int i = 0; i < 8; ++i)
				for (int j = 0; j < 8; ++j)
					if (owner[i][j] != ((i < mv && j < nv) ? 1 : 0))
						++bad;
			// optimal among the admissible shapes
			int best = ((mv + 1) / 2) * ((nv + 1) / 2);
			if (mv <= 4 && mv * ((nv + 3) / 4) < best) best = mv * ((nv + 3) / 4);
			if (nv <= 4 && ((mv + 3) / 4) * nv < best) best = ((mv + 3) / 4) * nv;
			if (am * an != best)
				++bad;
			std::printf("%d %d -> %dx%d grid, %dx%d atoms\n", mv, nv, gm, gn, am, an);
		}
	std::printf("bad=%d\n", bad);
	return bad != 0;
}
"""


def test_gemm_warp_grid_exhaustive():
    cuda_inc = "/usr/local/cuda/include"
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "wg.cpp")
        open(src, "w").write(SRC)
        exe = os.path.join(td, "wg")
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "quantit_b200", "csrc"), "-I", cuda_inc, src,
                        "-o", exe], check=True, capture_output=True, text=True)
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-2000:]
        assert "bad=0" in out.stdout
        # a full tile is the 2 x 2 grid of 4 x 4 atoms; a one-atom-high strip goes to a row of four warps
        assert "8 8 -> 2x2 grid, 4x4 atoms" in out.stdout
        assert "1 8 -> 1x4 grid, 1x2 atoms" in out.stdout
        assert "8 1 -> 4x1 grid, 2x1 atoms" in out.stdout


def _parallel_last_index(v, tol, pw, min_size, max_size):
    """the truncation rule as svd_select_kernel (quantit_b200/csrc/qtb_svd.cu) evaluates it: suffix sums S_i of |v|^pow,
    the largest index i >= min_size with S_i > tol^pow and i < max_size (max_size < 0: no cap); the minimum - 1 when no
    index qualifies; n - 1 when the reference's loop would not run at all (min_size > n - 1)."""
    import numpy as np

    n = len(v)
    if min_size > n - 1:
        return n - 1
    s = np.cumsum((np.abs(v) ** pw)[::-1])[::-1]
    ok = [i for i in range(min_size, n) if s[i] > tol ** pw and (max_size < 0 or i < max_size)]
    return max(ok) if ok else min_size - 1


def test_device_select_rule_equals_reference_loop():
    """compute_last_index (reference sources/LinearAlgebra.cpp:57-75, restated as a loop in the oracle) against the
    closed form the device kernel uses, on inputs whose sums are exact in fp64 (powers of two) so that the comparison
    with tol^pow cannot depend on the summation order: ties, zeros, every min / max combination, tol = 0."""
    import sys

    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qtb_oracle as orc

    rng = np.random.default_rng(0)
    checked = 0
    for _ in range(3000):
        n = int(rng.integers(1, 24))
        v = np.sort(2.0 ** -rng.integers(0, 12, n).astype(float))[::-1]
        if rng.random() < 0.3:
            v[int(rng.integers(0, n)):] = 0.0  # exact zeros at the tail (rank-deficient groups)
        tol = 0.0 if rng.random() < 0.2 else 2.0 ** -float(rng.integers(0, 14))
        min_size = int(rng.integers(0, n + 3))
        max_size = int(rng.integers(1, n + 3))
        want = orc.compute_last_index(v, tol, 2.0, min_size, max_size)
        got = _parallel_last_index(v, tol, 2.0, min_size, max_size)
        assert got == want, (v.tolist(), tol, min_size, max_size, got, want)
        checked += 1
    assert checked == 3000
