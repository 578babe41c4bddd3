"""world_size-2 gloo test of the N>1 plumbing (source-state broadcast + frame sharding), no GPU needed."""
import os
import sys
import pathlib

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from emoportraits_b200.config import shipped_config
    from emoportraits_b200.dist import broadcast_source_state, shard_frames, source_state_shapes

    cfg = shipped_config(256)
    st = None
    if rank == 0:
        g = torch.Generator().manual_seed(0)
        st = SimpleNamespace(**{k: torch.randn(s, generator=g) for k, s in source_state_shapes(cfg).items()})
    out = broadcast_source_state(st, cfg, "cpu", src=0)
    g = torch.Generator().manual_seed(0)
    ok = all(torch.equal(getattr(out, k), torch.randn(s, generator=g)) for k, s in source_state_shapes(cfg).items())
    ok = ok and torch.equal(out.source_theta_dev, out.pred_source_theta[0])
    q.put((rank, ok, shard_frames(7, rank, world)))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == [0, 2, 4, 6] and res[1][2] == [1, 3, 5]


def test_abi_library_exports_every_declared_symbol():
    """-m 'not gpu': the C-ABI library loads and exports every symbol include/emoportraits_b200.h declares"""
    import re
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as ge
    ge.build()
    from emoportraits_b200 import lib as L

    lib = L.load()
    header = (ROOT / "include" / "emoportraits_b200.h").read_text()
    declared = set(re.findall(r"\b(emo_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.emo_version() >= 100
