"""Host-side .pre reader/writer (bwbble_b200/csrc/precalc_io.cpp) against the file layout the reference writes
(store_sa_interval_list / precalc_sa_intervals, align.c:144-152,200-224): 4^12 records {int32 n; n x (u64 L, u64 U)}.
Compiled stand-alone with g++ (pure host code), so the format is checked without a GPU."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bwbble_b200", "csrc")

DRIVER = r"""
#include <cstdio>
#include <cstdlib>
#include "host_common.h"
int main(int argc, char **argv) {
    std::vector<uint32_t> sizes; std::vector<uint64_t> lu;
    if (bwb_host::read_pre_file(argv[1], sizes, lu)) return 2;
    unsigned long long total = 0;
    for (uint32_t s : sizes) total += s;
    printf("%zu %llu %zu\n", sizes.size(), total, lu.size());
    if (bwb_host::write_pre_file(argv[2], sizes, lu)) return 3;
    return 0;
}
"""


def test_pre_file_roundtrip_and_layout(tmp_path):
    exe = str(tmp_path / "pre_io")
    src = tmp_path / "driver.cpp"
    src.write_text(DRIVER)
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + CSRC, "-I" + os.path.join(ROOT, "include"), str(src),
                    os.path.join(CSRC, "precalc_io.cpp"), "-o", exe], check=True)
    rng = np.random.default_rng(3)
    rows = 1 << 24
    sizes = np.zeros(rows, dtype=np.int32)
    hot = rng.choice(rows, size=5000, replace=False)
    sizes[hot] = rng.integers(1, 6, size=len(hot))
    sizes[0], sizes[rows - 1] = 2, 3                     # first and last row populated
    total = int(sizes.sum())
    lu = rng.integers(0, 1 << 40, size=2 * total, dtype=np.uint64)
    # the reference's layout, record by record
    rec = np.zeros(rows * 4 + total * 16, dtype=np.uint8)
    starts = np.arange(rows, dtype=np.int64) * 4 + np.concatenate([[0], np.cumsum(sizes[:-1].astype(np.int64))]) * 16
    rec_view = rec
    sz_bytes = sizes.view(np.uint8).reshape(rows, 4)
    for k in range(4):
        rec_view[starts + k] = sz_bytes[:, k]
    lu_bytes = lu.view(np.uint8)
    pos = 0
    for r in np.nonzero(sizes)[0]:
        n = int(sizes[r])
        s = int(starts[r]) + 4
        rec_view[s:s + 16 * n] = lu_bytes[pos:pos + 16 * n]
        pos += 16 * n
    a, b = str(tmp_path / "a.pre"), str(tmp_path / "b.pre")
    rec.tofile(a)
    out = subprocess.run([exe, a, b], check=True, capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [rows, total, 2 * total]
    assert open(a, "rb").read() == open(b, "rb").read()
    # a truncated file is an error, not a short table
    open(a, "r+b").truncate(os.path.getsize(a) - 8)
    assert subprocess.run([exe, a, b]).returncode == 2
