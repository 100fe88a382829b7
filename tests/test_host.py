"""CPU: host-side logic — state-dict key grammar, batch sharding, "no CPU path" behaviour, and the N>1 shard
bookkeeping with a world_size-2 gloo group (the data path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def test_state_dict_keys_match_paddle_grammar():
    """226 tensors; Sequential children by index; BN keys weight/bias/_mean/_variance."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import LWSNet
    m = LWSNet(O.default_args())
    sd = m.state_dict()
    assert len(sd) == 226
    assert set(sd) == set(O.build_oracle(0).state_dict())
    for k in ("feature_extraction.dres0.0.0.weight", "feature_extraction.dres0.0.1._variance",
              "feature_extraction.dres2.conv5.0.weight", "feature_extraction.classif1.2.weight",
              "volume_postprocess.0.1.0._mean", "volume_postprocess.0.1.2.weight", "volume_postprocess.2.5.2.weight",
              "refinement1_left.0.weight", "refinement1_disp.4.3.weight", "refinement2.0.0.weight", "refinement2.0.2.weight",
              "refinement2.5.weight"):
        assert k in sd, k
    assert tuple(sd["volume_postprocess.0.1.2.weight"].shape) == (32, 32, 3, 3, 3)
    assert tuple(sd["volume_postprocess.1.0.2.weight"].shape) == (8, 1, 3, 3, 3)
    assert tuple(sd["feature_extraction.dres2.conv6.0.weight"].shape) == (16, 8, 3, 3)
    assert tuple(sd["refinement1_left.2.2.weight"].shape) == (32, 1, 3, 3)
    assert sum(p.numel() for p in m.parameters()) == 177890


def test_reference_args_are_honoured():
    from types import SimpleNamespace
    from lwsnet_b200 import LWSNet
    m = LWSNet(SimpleNamespace(maxdisplist=[48, 5, 5], layers_3d=2, channels_3d=4, growth_rate=[4, 2, 2]))
    assert m.maxdisplist == [48, 5, 5] and len(m.volume_postprocess) == 3
    assert len(m.volume_postprocess[0]) == 4 and m.volume_postprocess[0].channels == 16 and m.volume_postprocess[1].channels == 8


def test_no_cpu_path():
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import LWSNet, disparity_regression, ops
    from lwsnet_b200._lib import LwsError
    m = LWSNet(O.default_args())
    x = torch.zeros(1, 3, 64, 128)
    with pytest.raises(LwsError):
        m(x, x)
    with pytest.raises(LwsError):
        m.warp(torch.zeros(1, 4, 8, 8), torch.zeros(1, 1, 8, 8))
    with pytest.raises(LwsError):
        m._build_volume_2d(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8), 4)
    with pytest.raises(AssertionError):  # reference assert, models/models.py:63
        m._build_volume_2d(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8), 5, stride=2)
    with pytest.raises(LwsError):
        m.volume_postprocess[1](torch.zeros(1, 1, 9, 8, 8))
    with pytest.raises(LwsError):
        disparity_regression(0, 4)(torch.full((1, 4, 2, 2), 0.25))
    with pytest.raises(LwsError):
        ops.softmax_regression(torch.zeros(1, 4, 2, 2), 0.0)


def test_product_does_not_import_oracle():
    import ast
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lwsnet_b200")
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            tree = ast.parse(open(os.path.join(root, fn)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), f"{fn} imports oracle"


def test_shard_range_and_micro_batches():
    from lwsnet_b200.runner import micro_batches, shard_range
    for total in (256, 64, 32, 7, 1, 0):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(256, 3, 8) == (96, 128)
    assert micro_batches(5, 2) == [(0, 2), (2, 4), (4, 5)]
    from lwsnet_b200.runner import ramp_batches
    assert ramp_batches(64, 32, 8) == [(0, 8), (8, 40), (40, 56), (56, 64)]
    assert ramp_batches(10, 24, 8) == [(0, 8), (8, 10)] and ramp_batches(3, 2, 2) == [(0, 2), (2, 3)]
    for n, mb, e in ((64, 24, 8), (17, 8, 8), (1, 4, 2), (100, 7, 3)):
        ch = ramp_batches(n, mb, e)
        assert ch[0][0] == 0 and ch[-1][1] == n and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert all(0 < hi - lo <= mb for lo, hi in ch)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, total, q):
    import torch.distributed as dist
    from lwsnet_b200.runner import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    # each rank "processes" its shard: here the identity on the pair indices; a gather (outside any timed region in
    # the real runner) must reassemble the batch in order, and the max-over-ranks timing reduction must agree
    mine = torch.arange(lo, hi, dtype=torch.float32)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    parts = [torch.zeros(int(s.item())) for s in sizes]
    dist.all_gather(parts, mine) if len({int(s.item()) for s in sizes}) == 1 else None
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, lo, hi, [p.tolist() for p in parts], t.item()))
    dist.destroy_process_group()


def test_two_rank_shard_bookkeeping_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 8)]
    for r in res:
        assert sum(r[3], []) == [float(i) for i in range(8)]
        assert r[4] == 11.0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and d["n_gpus"] == 1 and d["steps"] == 1
