"""CPU: the C-ABI library loads and exports every symbol include/klsh.h declares, fails loudly
without a GPU, and the host-side pieces (CLI, log parsing, ctypes mirror) behave."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import kmerlsh_b200
from kmerlsh_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "klsh.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(klsh_[a-z_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(api.lib_path())
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(s[0] for s in api.SYMBOLS) == names  # the ctypes mirror binds all of them


def test_no_cpu_fallback():
    """Without a usable B200 the product path must fail loudly, never compute."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(kmerlsh_b200.KlshError, match="no CUDA device|no CPU fallback"):
        kmerlsh_b200.Context(0)
    with pytest.raises(kmerlsh_b200.KlshError):
        kmerlsh_b200.Cluster((np.zeros((4, 2), np.float32), None, None), 0.8, 1, 1, 2, 10)


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kmerlsh_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "klsh_oracle" not in text and "oracle_lib" not in text and "klo_" not in text, f
                assert "import oracle" not in text and "from oracle" not in text, f
    ldd = subprocess.run(["ldd", api.lib_path()], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in ldd


def test_cli_rejects_other_modes(tmp_path):
    exe = os.path.join(ROOT, "kmerlsh_b200", "kmerLSH_b200")
    assert os.path.exists(exe)
    for mode in ("K", "B", ""):
        r = subprocess.run([exe, "-M", mode] if mode else [exe], cwd=str(tmp_path), stderr=subprocess.PIPE, text=True)
        assert r.returncode == 2 and "modes C and E" in r.stderr, mode


def test_log_line_roundtrip():
    cov = np.array([123456.789012, 0.5, 99999999.25])
    line = synth.format_log_line(1000, cov)
    assert line == "1000\t123456.789012\t0.500000\t99999999.250000\n"
    kmap, c32 = synth.parse_log_line(line, 3)
    assert kmap == 1000 and c32.dtype == np.float32
    vk = synth.v_kmers_from_cov(c32, kmap)
    assert vk[1] == np.float32(0.5) / np.float32(1000)


def test_floor_log2_matches_reference_formula():
    """H = floor(log2(N)) (function/cluster.cc:194): the integer form the library uses agrees with
    the double form at every power-of-two edge up to 2^40."""
    for k in range(1, 41):
        for n in (2 ** k - 1, 2 ** k, 2 ** k + 1):
            assert int(np.floor(np.log2(float(n)))) == n.bit_length() - 1


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm) needs no GPU: it times the reference's own
    binary on a bounded sample and prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    ref_bin = os.path.join(ROOT, "oracle", "_ref", "kmerLSH_ref")
    if not os.path.exists(ref_bin):
        pytest.skip("oracle/_ref/kmerLSH_ref not built (needs /root/reference)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-rows", "30000", "--cpu-sample-iters", "3"], check=True,
                         capture_output=True, text=True, cwd=ROOT).stdout.strip().splitlines()[-1]
    line = json.loads(out)
    assert line["impl"] == "reference" and line["unit"] == "rows/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["workload"].startswith("C1")
