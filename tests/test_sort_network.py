"""The sorting networks compiled into k_step_fast (mettagrid_b200/csrc/mg_fast.cu): the pair lists in the source
must sort every 0-1 input (0-1 principle), i.e. every input."""

import re
from pathlib import Path

SRC = Path(__file__).resolve().parent.parent / "mettagrid_b200" / "csrc" / "mg_fast.cu"


def _pairs(n):
    text = SRC.read_text()
    m = re.search(r"/\* SORT%d \*/(.*?)\n\}" % n, text, re.S)
    assert m, f"no network for {n} keys"
    return [(int(a), int(b)) for a, b in re.findall(r"CE\((\d+), (\d+)\)", m.group(1))]


def _sorts_all_01(n, pairs):
    for bits in range(1 << n):
        v = bits
        for a, b in pairs:  # exchange so that position a holds the smaller bit
            if (v >> a) & 1 and not (v >> b) & 1:
                v ^= (1 << a) | (1 << b)
        ones = bin(v).count("1")
        if v != ((1 << ones) - 1) << (n - ones):
            return False
    return True


def test_networks_sort():
    for n in (8, 16):
        pairs = _pairs(n)
        assert len(pairs) == {8: 19, 16: 63}[n]
        assert all(0 <= a < b < n for a, b in pairs)
        assert _sorts_all_01(n, pairs)
