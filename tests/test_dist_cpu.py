"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: sigma sharding, the final host gather
merged in sigma order, and the max/sum timing reductions bench.py uses.  The per-shard compute is
stood in for by the oracle (allowed here: tests only)."""
import os, sys, json, subprocess, tempfile, textwrap
import pytest
from conftest import ROOT, GOLDEN

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import torch.distributed as dist
    import oracle_lib as O
    from avx_ecm_b200 import dist as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
    g = json.load(open(%(golden)r))
    N, b1, total, s0 = int(g["n"]), g["b1"], len(g["save_lines"]), int(g["sigma0"])
    def compute(first_sigma, count):
        rs = [O.ecm_curve(N, b1, b1, first_sigma + i) for i in range(count)]
        return {"save_lines": [r["save_line"] for r in rs], "factors": [(first_sigma + i, 1, r["f1"]) for i, r in enumerate(rs) if r["f1"]]}
    merged = D.run_sharded(total, s0, compute)
    tmax = D.all_max(1.0 + dist.get_rank())
    tsum = D.all_sum(10.0 * (1 + dist.get_rank()))
    if dist.get_rank() == 0:
        json.dump({"merged": merged, "tmax": tmax, "tsum": tsum}, open(sys.argv[2], "w"))
    dist.destroy_process_group()
''')


def test_shard_range_partitions_exactly():
    from avx_ecm_b200.dist import shard_range
    for total in (1, 7, 8, 65536, 1000003):
        for world in (1, 2, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


@pytest.mark.parametrize("name", ["syn415_b1_3e4_s1only", "csh250k_stage1_factor"])
def test_two_rank_gloo_gather_equals_reference_file(name):
    golden = os.path.join(ROOT, "tests", "golden", name + ".json")
    with tempfile.TemporaryDirectory() as d:
        script = os.path.join(d, "w.py")
        port = 29500 + os.getpid() % 2000
        open(script, "w").write(WORKER % {"root": ROOT, "port": port, "golden": golden})
        out = os.path.join(d, "out.json")
        procs = [subprocess.Popen([sys.executable, script, str(r), out]) for r in range(2)]
        for p in procs:
            assert p.wait(timeout=600) == 0
        res = json.load(open(out))
    g = GOLDEN[name]
    # merged output of the two ranks == the reference's save_b1.txt (threads=1), byte for byte
    assert res["merged"]["save_lines"] == g["save_lines"]
    assert res["merged"]["sigmas"] == [int(g["sigma0"]) + i for i in range(len(g["save_lines"]))]
    assert [[int(f["sigma"]), f["stage"], int(f["factor"])] for f in g["factors"]] == [[s, st, f] for s, st, f in res["merged"]["factors"]]
    assert res["tmax"] == 2.0 and res["tsum"] == 30.0
