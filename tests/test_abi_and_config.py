"""CPU tests: the C-ABI library loads and exports every declared symbol; MOR_config.txt grammar."""
import ctypes as C
import re

import pytest

from dynamicslamtool_b200.binding import MorConfig, PRODUCT_LIB
from helpers import DEFAULT_CFG, ROOT, write_cfg


def declared_functions():
    text = (ROOT / "include" / "mor_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mor_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(built):
    lib = C.CDLL(str(PRODUCT_LIB))
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/mor_b200.h but not exported: {missing}"


def test_oracle_exports_the_shared_subset(oracle):
    for n in ("create", "create_ex", "destroy", "push_raw_cloud_and_pose", "filter_cloud", "tap", "parse_config", "get_config"):
        assert hasattr(oracle.lib, "oracle_" + n)


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_default_config_values(which, product, oracle):
    b = product if which == "product" else oracle
    cfg = MorConfig()
    assert b.parse_config(str(DEFAULT_CFG).encode(), C.byref(cfg)) == 0
    # the 23 upstream defaults (reference config/MOR_config.txt:1-39)
    assert cfg.method_choice == 2 and cfg.opc_normalization_factor == 20
    assert cfg.min_cluster_size == 200 and cfg.max_cluster_size == 35000
    assert abs(cfg.ec_distance_threshold - 0.11) < 1e-7 and abs(cfg.gp_limit + 0.5) < 1e-7
    assert (cfg.trim_x, cfg.trim_y, cfg.trim_z) == (3.0, 3.0, 5.0)
    assert abs(cfg.pde_lb - 0.005) < 1e-8 and abs(cfg.pde_ub - 0.5) < 1e-7 and abs(cfg.pde_distance_threshold - 0.15) < 1e-7
    assert abs(cfg.volume_constraint - 0.3) < 1e-7 and abs(cfg.leave_off_distance - 0.5) < 1e-7 and abs(cfg.catch_up_distance - 0.3) < 1e-7
    assert abs(cfg.gp_leaf - 0.1) < 1e-7 and cfg.bin_gap == 10.0
    assert cfg.output_topic == b"/output" and cfg.input_odometry_topic == b"/camera/odom/sample" and cfg.debug_fid == b"/debug"
    assert cfg.ground_mode == 0


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_config_errors_and_grammar(which, product, oracle, tmp_path):
    b = product if which == "product" else oracle
    cfg = MorConfig()

    def parse(p):
        return b.parse_config(str(p).encode(), C.byref(cfg))

    assert parse(tmp_path / "nope.txt") == 1                                        # cpp:703-707
    assert parse(write_cfg(tmp_path, "a.txt", extra=["bogus_key:1"])) == 2          # cpp:856-860
    assert parse(write_cfg(tmp_path, "b.txt", trim_x="abc")) == 3                   # std::stof throws
    assert parse(write_cfg(tmp_path, "c.txt", drop=("pde_ub",))) == 4               # uninitialised member in the reference
    assert parse(write_cfg(tmp_path, "d.txt", method_choice=3)) == 3                # cpp:568-593 UB
    # every ':' after the first is dropped from the value (cpp:718-733); '#' lines and short lines are skipped
    p = write_cfg(tmp_path, "e.txt", output_topic="/a:b:c", extra=["#comment:1", "ab", ""])
    assert parse(p) == 0 and cfg.output_topic == b"/abc"
    # opc_normalization_factor goes through stof into an int (cpp:843)
    assert parse(write_cfg(tmp_path, "f.txt", opc_normalization_factor="7.9")) == 0 and cfg.opc_normalization_factor == 7
    # extension keys
    assert parse(write_cfg(tmp_path, "g.txt", ground_mode=2, gp_planarity=0.02)) == 0 and cfg.ground_mode == 2
    assert parse(write_cfg(tmp_path, "h.txt", ground_mode=5)) == 3


def test_product_fails_loudly_without_a_device(product):
    """On the CPU box there is no device: creation must fail with MOR_ERR_CUDA, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    st = product.create_ex(str(DEFAULT_CFG).encode(), 4, 3, 0, None, C.byref(h))
    assert st == 7 and not h.value


@pytest.mark.parametrize("seed", range(80))
def test_config_parser_differential_fuzz(seed, product, oracle, tmp_path):
    """The product's parser (csrc/mor_config.cpp) and the oracle's are two independent restatements of setVariables
    (cpp:698-864). Random mutations of the default file - shuffled lines, comments, short lines, extra colons, spaces,
    odd numbers, dropped / duplicated / unknown keys - must give the same status and, on success, the same struct."""
    import random
    rnd = random.Random(seed)
    lines = [l for l in DEFAULT_CFG.read_text().splitlines() if len(l) >= 3 and not l.startswith("#")]
    rnd.shuffle(lines)
    out = []
    for l in lines:
        key, val = l.split(":", 1)
        r = rnd.random()
        if r < 0.01:
            continue                                                     # dropped key
        if r < 0.05:
            out.append(l)                                                # duplicated key (the later line wins)
        if r < 0.10:
            val = rnd.choice(["1e-1", "007", "+3", "-0.5", "2.", ".5", "0x10", "1,5", "nan", "abc", "", " 4", "4 ", "3:3", "1e400"])
        if r > 0.99:
            key = key + "_x"                                             # unknown key
        out.append(f"{key}:{val}")
        if rnd.random() < 0.2:
            out.append(rnd.choice(["# note: with colon", "#", "ab", "", "  ", "#x", "a:" if rnd.random() < 0.1 else "# a:"]))
    p = tmp_path / "fuzz.txt"
    p.write_text("\n".join(out) + ("\n" if rnd.random() < 0.5 else ""))
    ca, cb = MorConfig(), MorConfig()
    sa, sb = product.parse_config(str(p).encode(), C.byref(ca)), oracle.parse_config(str(p).encode(), C.byref(cb))
    assert sa == sb, f"status product {sa} vs oracle {sb} for\n{p.read_text()}"
    if sa == 0:
        assert bytes(ca) == bytes(cb), p.read_text()


def test_stream_depth_constant_is_what_bench_and_tests_assume():
    """include/mor_b200.h: MOR_STREAM_DEPTH frames may be in flight; bench.py and the streaming test keep that many buffers."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    depth = int(re.search(r"#define MOR_STREAM_DEPTH (\d+)", (root / "include" / "mor_b200.h").read_text()).group(1))
    assert depth == int(re.search(r"DEPTH = (\d+)  # MOR_STREAM_DEPTH", (root / "bench.py").read_text()).group(1))
    assert depth == int(re.search(r"DEPTH = (\d+)  # MOR_STREAM_DEPTH", (root / "tests" / "test_gpu_streaming.py").read_text()).group(1))
