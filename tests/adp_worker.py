"""Worker of tests/test_gpu_parity.py::test_cuda_adp_matches_reference_fixture (own process: device code that has
never run on a GPU must not be able to take the CUDA context of the other tests down with it)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common  # noqa: E402
from imd_b200 import api  # noqa: E402


def main(name, tmp, lanes):
    g = common.load_golden(name)
    sim = common.make_sim(api.IMDB200, g, tmp, lanes_per_atom=int(lanes))
    out = common.run_protocol(sim, g)
    errs = common.compare(out, g, full_list=True, rtol=1e-10, traj_rtol=1e-8)
    assert "f0:adp_mu" in errs and "f0:adp_lambda" in errs
    sim.close()
    print("ADP_OK", name, lanes, {k: f"{v:.1e}" for k, v in errs.items()})


if __name__ == "__main__":
    main(*sys.argv[1:4])
