"""bench.py's JSON assembly on CPU: run_b200 over the numpy test double of the device handle (tests/fake_device.py) with
made-up timings -- guards the contract keys of the bench line for both TRSM paths.  Nothing here measures anything."""
import json
import os
import sys
import types

import pytest

from fake_device import FakeHandle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KEYS = ["kmat_ms", "chol_ms", "solve_ms", "kstar_ms", "trsm_ms", "grad_ms", "n_trsm", "n_launches", "fit_ms",
        "predict_device_wall_ms", "predict_d2h_wall_ms", "i8_prep_ms", "i8_check_ms", "i8_rows_ms", "i8_block_rows", "i8_fallbacks"]


@pytest.mark.parametrize("i8", [True, False])
def test_bench_line_contract(monkeypatch, capsys, i8):
    import bench
    from mogp_emulator_b200 import libmogp

    class Handle(FakeHandle):
        def timings(self, reset=False):
            d = {k: 1.0 for k in KEYS}
            d["i8_block_rows"] = 2.0 if i8 else 0.0
            d["i8_fallbacks"] = 0.0
            return d

    monkeypatch.setattr(libmogp, "Handle", Handle)
    monkeypatch.setattr(libmogp, "HAVE_LIBMOGP", True)
    monkeypatch.setattr(libmogp, "gpu_usable", lambda: True)
    monkeypatch.setattr(libmogp, "peak_dmma_tflops", lambda device=0, iters=0: 37.0)
    monkeypatch.setattr(libmogp, "peak_i8_tops", lambda device=0, iters=0: (4500.0, 3000.0))
    for var in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="b200", workload="c1", no_cpu=False, no_e2e=False, no_other=True)
    bench.run_b200(args, bench.WORKLOADS["c1"])
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["steps"] == 2 and line["n_gpus"] == 1 and line["higher_is_better"] is False
    assert line["dtype"].startswith("f64") and (("int8" in line["dtype"]) == i8)      # the arithmetic of the TRSM is disclosed
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roof, key
    assert roof["bound"] == "tensor" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert ("i8_trsm_kernel" in roof["kernel"]) == i8 and roof["unit"] == ("TOP/s" if i8 else "TFLOP/s")
    assert line["config"]["trsm_path"].startswith("int8 tcgen05" if i8 else "FP64 DMMA")
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert set(("output", "owner_rank", "mean_max_rel", "var_max_abs", "var_worst_vs_tolerance")) <= set(line["parity_vs_cpu_sample"])
