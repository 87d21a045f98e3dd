"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/imd_b200.h declares;
its host-side helpers (no CUDA) behave like the reference's readers.  No compute entry point is called."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import common

ROOT = common.ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "imd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(imdb200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built_lib):
    from imd_b200 import api
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(built_lib, s)]
    assert not missing, missing
    assert sorted(api.EXPORTS) == syms, "imd_b200/api.py binds a different set than the header declares"


def test_no_gpu_means_loud_failure(built_lib):
    """Without a device the product must refuse to run rather than fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from imd_b200 import api
    with pytest.raises(api.IMDError):
        api.IMDB200(1, np.eye(3) * 20.0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "imd_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "imd_oracle" not in txt \
                    and "liboracle" not in txt, f


def test_host_table_reader_matches_oracle_reader(built_lib, tmp_path):
    """imdb200_read_pot_table (host C) against the oracle's restatement of read_pot_table on both formats."""
    from imd_b200 import api
    from oracle import oracle as orc
    for name, nt in (("nial_nvt", 2), ("lj_nve", 2), ("cu_nve", 1)):
        g = common.load_golden(name)
        paths = common.write_tables(g, str(tmp_path / name))
        o = orc.OracleIMD(nt, np.eye(3) * 30.0, pair=paths["pair"], embed=paths.get("embed"), rho=paths.get("rho"))
        eam = "rho" in paths
        for which, key, ncols, radial in ((0, "pair", nt * nt, 1), (1, "embed", nt, 0), (2, "rho", nt * nt, 1)):
            if key not in paths:
                continue
            pt, cellsz = api.read_pot_table(paths[key], ncols, radial, nt, 2 if eam else 1)
            for col in range(ncols):
                info = o.table_info(which, col)
                assert pt.begin[col] == info["begin"] and pt.end[col] == info["end"]
                assert pt.step[col] == info["step"] and pt.len[col] == info["len"]
                # table rows incl. the two extrapolated pad rows, probed through PAIR_INT at the nodes
                n = pt.len[col]
                tab = np.array([pt.table[k * ncols + col] for k in range(n + 2)])
                x = info["begin"] + info["step"] * np.arange(n)
                v, _ = o.pair_int(which, col, x)
                assert np.allclose(v[:-1], tab[:n - 1], rtol=0, atol=1e-12 * max(1.0, np.abs(tab).max()))
            built_lib.imdb200_free_pot_table(C.byref(pt))


def test_ctypes_mirror_has_the_layout_of_the_c_structs(tmp_path):
    """imdb200_config / imdb200_scalars as gcc lays them out against imd_b200/api.py's ctypes mirror: sizes and the
    offsets of the fields behind the first alignment gap (a silent mismatch would scramble every parameter)."""
    import ctypes as C
    import subprocess
    from imd_b200 import api
    src = tmp_path / "layout.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "imd_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(imdb200_config), offsetof(imdb200_config, nbl_margin),
         offsetof(imdb200_config, ensemble), offsetof(imdb200_config, interpolation), offsetof(imdb200_config, xi),
         sizeof(imdb200_scalars), offsetof(imdb200_scalars, cellsz));
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(common.ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(api.Config), api.Config.nbl_margin.offset, api.Config.ensemble.offset, api.Config.interpolation.offset,
            api.Config.xi.offset, C.sizeof(api.Scalars), api.Scalars.cellsz.offset]
    assert got == want, (got, want)
