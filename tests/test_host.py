"""CPU-side checks: the C-ABI library loads and exports every symbol include/eolc.h declares, the product fails loudly
without a GPU (no CPU fallback), host helpers (mesh generators, ArcSim edge order), and the world-size-2 sharding of
the ensemble driver (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import eol_cloth_b200 as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "eolc.h")).read()
    names = sorted(set(re.findall(r"\b(eolc_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = ctypes.CDLL(E.lib_path())
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/eolc.h but not exported"


def test_contact_struct_layout():
    from oracle import oracle as O
    assert E.CONTACT_DTYPE == O.CONTACT_DTYPE and E.CONTACT_DTYPE.itemsize == 264


def test_no_cpu_fallback():
    if E.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(E.EolcError, match="no CPU fallback"):
        E.Context(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "eol_cloth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_regular2_and_build4_counts():
    X, fn = E.meshgen.regular2(64)
    assert X.shape == (4096, 2) and fn.shape == (7938, 3)
    es = E.meshgen.edge_stencils(4096, fn)
    assert es.shape[0] == 3 * 63 * 63 + 2 * 63 and int((es[:, 3] >= 0).sum()) == 11781
    X, fn = E.meshgen.build4(512)
    assert X.shape[0] == 523265 and fn.shape[0] == 1044484
    # positive orientation everywhere
    for gen in (E.meshgen.regular2, E.meshgen.build4):
        X, fn = gen(5)
        a, b, c = X[fn[:, 0]], X[fn[:, 1]], X[fn[:, 2]]
        area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
        assert np.all(area2 > 0)


def test_edge_stencils_follow_arcsim_rule():
    """Pure-Python restatement of Mesh::add(Face) (mesh.cpp:356-378) on a small mesh."""
    X, fn = E.meshgen.build4(4)
    edges, index = [], {}
    for face in fn.tolist():
        for i in range(3):
            a, b = face[i], face[(i + 1) % 3]
            if (min(a, b), max(a, b)) not in index:
                index[(min(a, b), max(a, b))] = len(edges)
                edges.append([a, b, -1, -1])
        for i in range(3):
            v0, v1 = face[(i + 1) % 3], face[(i + 2) % 3]
            e = edges[index[(min(v0, v1), max(v0, v1))]]
            e[2 + (0 if e[0] == v0 else 1)] = face[i]
    assert np.array_equal(E.meshgen.edge_stencils(X.shape[0], fn), np.array(edges, np.int32))


def test_bench_sharding_world_size_2_gloo(tmp_path):
    """bench.py's ensemble partition + max-over-ranks reduction on 2 CPU ranks (gloo)."""
    script = tmp_path / "shard.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "lo, hi = bench.shard_range(4096, r, w)\n"
        "t = bench.max_over_ranks(float(r + 1), 'cpu')\n"
        "tot = bench.sum_over_ranks(float(hi - lo), 'cpu')\n"
        "assert t == 2.0 and tot == 4096.0, (t, tot)\n"
        "assert (lo, hi) == ((0, 2048) if r == 0 else (2048, 4096))\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)], env=env, timeout=300)


def test_reference_arm_prints_one_line_under_torchrun():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 alone runs the CPU path and prints ONE JSON line with the reference-arm
    keys; the other rank exits 0 without work."""
    import json
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29518", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "elements/s" and d["higher_is_better"] is True and d["value"] > 0
    cb = d["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libforces_ref.so"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["single_thread_value"] > 0
    assert cb["value"] == d["value"] and cb["port_threaded"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
