"""The C-ABI library loads and exports every symbol include/incerto_b200.h declares (no compute here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "incerto_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(incerto_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(engine_lib):
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(engine_lib, n), f"{n} declared in include/incerto_b200.h but not exported"


def test_ffi_symbol_list_matches_header():
    from incerto_b200 import _ffi
    assert sorted(_ffi.SYMBOLS) == _declared_symbols()


def test_abi_version(engine_lib):
    assert engine_lib.incerto_abi_version() == 1


def test_enum_values_match_header():
    from incerto_b200 import _ffi
    text = open(os.path.join(ROOT, "include", "incerto_b200.h")).read()
    for name, val in re.findall(r"\b(INCERTO_[A-Z0-9_]+)\s*=\s*(\d+)", text):
        short = name[len("INCERTO_"):]
        if hasattr(_ffi, short):
            assert getattr(_ffi, short) == int(val), name


def test_param_struct_sizes(engine_lib):
    """ctypes mirrors must have the C layout (checked against sizes computed by the C compiler)."""
    import subprocess, tempfile
    from incerto_b200 import _ffi
    pairs = [("incerto_add_const_params", _ffi.AddConstParams), ("incerto_add_column_params", _ffi.AddColumnParams),
             ("incerto_coin_toss_params", _ffi.CoinTossParams), ("incerto_forest_params", _ffi.ForestParams),
             ("incerto_spawn_forest_params", _ffi.SpawnForestParams), ("incerto_pandemic_params", _ffi.PandemicParams),
             ("incerto_spawn_pandemic_params", _ffi.SpawnPandemicParams), ("incerto_air_params", _ffi.AirParams),
             ("incerto_spawn_aircraft_params", _ffi.SpawnAircraftParams),
             ("incerto_aggregate_desc", _ffi.AggregateDesc), ("incerto_filter", _ffi.Filter),
             ("incerto_fire_stats", _ffi.FireStats), ("incerto_pandemic_stats", _ffi.PandemicStats),
             ("incerto_kernel_stat", _ffi.KernelStat), ("incerto_system_desc", _ffi.SystemDesc),
             ("incerto_column_data", _ffi.ColumnData),
             ("incerto_spawn_coin_tossers_params", _ffi.SpawnCoinTossersParams),
             ("incerto_spatial_params", _ffi.SpatialParams), ("incerto_spawn_spatial_params", _ffi.SpawnSpatialParams),
             ("incerto_traders_params", _ffi.TradersParams), ("incerto_spawn_traders_params", _ffi.SpawnTradersParams),
             ("incerto_spawn_stocks_params", _ffi.SpawnStocksParams)]
    src = '#include <stdio.h>\n#include "incerto_b200.h"\nint main(void){' + "".join(
        f'printf("%zu\\n", sizeof({c}));' for c, _ in pairs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")],
                       check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    for (cname, ct), size in zip(pairs, out):
        assert C.sizeof(ct) == int(size), cname


def test_builder_calls_without_device(engine_lib):
    """Builder bookkeeping needs no GPU; `build` must fail loudly (no CPU fallback) when there is none."""
    import incerto_b200 as ib
    from incerto_b200 import _ffi as F
    b = ib.SimulationBuilder(seed=1)
    c = b.component("MyValue", F.U64)
    m = b.component("MyMarker")
    # tests/test_builder.rs:27-42
    b.record_aggregate_time_series(ib.Sum(c), 8)
    with pytest.raises(ib.BuilderError) as e:
        b.record_aggregate_time_series(ib.Sum(c), 12)
    assert e.value.kind == "TimeSeriesRecordingConflict"
    # :44-57 different Out is fine; :59-72 different filter is fine
    b.record_aggregate_time_series(ib.Count(c), 8)
    b.record_aggregate_time_series(ib.Sum(c).filtered(ib.With(m)), 8)
    with pytest.raises(ib.EngineError):   # assert!(sample_interval > 0), simulation_builder.rs:271
        b.record_aggregate_time_series(ib.Minimum(c), 0)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(ib.EngineError) as e:
            b.build()
        assert e.value.code == F.ERR_NO_DEVICE


def test_rust_sys_crate_declares_every_symbol():
    """rust/incerto-cuda-sys cannot be built here (no cargo); at least its hand-written extern block must name exactly
    the functions of the header, and the safe façade must only call functions that block declares."""
    sys_rs = open(os.path.join(ROOT, "rust", "incerto-cuda-sys", "src", "lib.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (incerto_[a-z0-9_]+)\s*\(", sys_rs)))
    assert declared == _declared_symbols()
    facade = open(os.path.join(ROOT, "rust", "incerto-cuda", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(incerto_[a-z0-9_]+)\s*\(", facade))
    assert used and used <= set(declared), used - set(declared)
