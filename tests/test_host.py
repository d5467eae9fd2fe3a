"""CPU tests of the host-side logic: collate parity (bit exact), the C-ABI boundary (header <->
ctypes <-> exported symbols, no compute calls), loud failure without a GPU, DDP plumbing under
gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import make_batch, oracle_graphs
import oracle.data as od

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["MUTAG", "ZINC", "PATTERN", "CLUSTER", "MOLHIV"])
def test_collate_host_bit_exact_vs_reference_loops(name):
    ids = np.array([5, 0, 9, 3, 3, 7])                                   # unordered, with a repeat
    cfg, graphs, store, out = make_batch(name, 10, seed=7, ids=ids)
    ogs = oracle_graphs([graphs[i] for i in ids], cfg['n_tags'])
    fn = {'v2': od.collate_v2, 'sbm': od.collate_sbm, 'ogb': od.collate_v2}[cfg['kind']]
    ref = fn(ogs, n_tags=cfg['n_tags'], n_features=store.n_features)
    for i, (a, b) in enumerate(zip(out, ref)):
        assert (a is None) == (b is None), i
        if a is not None:
            assert a.shape == b.shape and torch.equal(a.to(b.dtype), b), (name, i)
    # integer outputs keep the reference dtypes
    assert out[1].dtype == torch.bool and out[6].dtype == torch.int64 and out[8].dtype == torch.int64


def test_collate_single_node_and_edgeless_graphs():
    from feta_tmlr_b200 import data as fdata
    graphs = [dict(x=np.zeros((1, 1), dtype=np.int64), edge_index=np.zeros((2, 0), dtype=np.int64),
                   y=np.int64(0), degree=np.ones(1, dtype=np.float32), pe=None, lap_pe=None),
              dict(x=np.ones((3, 1), dtype=np.int64), edge_index=np.array([[0, 1], [1, 0]]), y=np.int64(1),
                   degree=np.ones(3, dtype=np.float32), pe=None, lap_pe=None)]
    store = fdata.GraphStore(graphs, kind='v2', n_tags=2)
    px, mask, pe, lap, deg, y, ei, bi, fi = fdata.collate_host(store, [0, 1])
    assert mask.tolist() == [[False, True, True], [False, False, False]]
    assert ei.tolist() == [[1, 2], [2, 1]] and bi.tolist() == [0, 1, 1, 1]
    assert fi.tolist() == [[0, 0], [1, 0], [1, 1], [1, 2]]


@pytest.mark.parametrize("name", ["ZINC", "PATTERN"])
def test_collate_static_shapes_pad_with_ignored_columns(name):
    """static=(nmax_cap, e_cap): fixed shapes for CUDA-graph replays; the real part equals the reference tuple,
    edge padding is (-1, -1) (ignored by the plan builder), node-level labels pad with -100, and the int32 edge
    option carries the same values."""
    from feta_tmlr_b200 import data as fdata
    cfg, graphs, store, ref = make_batch(name, 6, seed=3)
    nmax = ref[0].shape[1]
    E = ref[6].shape[1]
    st = fdata.collate_host(store, np.arange(6), static=(nmax + 5, E + 37))
    st32 = fdata.collate_host(store, np.arange(6), static=(nmax + 5, E + 37), edge_dtype=np.int32)
    assert st[0].shape[1] == nmax + 5 and st[6].shape == (2, E + 37) and st[6].dtype == torch.int64
    assert torch.equal(st[0][:, :nmax], ref[0]) and bool((st[0][:, nmax:] == 0).all())
    assert torch.equal(st[1][:, :nmax], ref[1]) and bool(st[1][:, nmax:].all())
    assert torch.equal(st[6][:, :E], ref[6]) and bool((st[6][:, E:] == -1).all())
    assert st32[6].dtype == torch.int32 and torch.equal(st32[6].long(), st[6])
    if cfg['head'] == 'node':
        assert st[5].shape == (6, nmax + 5) and torch.equal(st[5][~st[1]], ref[5]) and bool((st[5][st[1]] == -100).all())
    with pytest.raises(ValueError):
        fdata.collate_host(store, np.arange(6), static=(nmax - 1, E))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "feta_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|int64_t|const char\*)\s+(feta_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_abi_header_ctypes_and_exports_agree():
    """Every entry point include/feta_b200.h declares is exported by the .so and bound with the
    right arity (no compute calls: this runs without a GPU)."""
    from feta_tmlr_b200 import _lib
    decl = _header_functions()
    assert len(decl) >= 25
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, n in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == n, (name, n, len(_lib.SIGNATURES[name][1]))
    lib = _lib.load()
    for name in decl:
        assert hasattr(lib, name)
    assert lib.feta_version() >= 100
    assert lib.feta_last_error_string() is not None
    assert lib.feta_cheb_plan_workspace_bytes(100, 1000) > 0 and lib.feta_cheb_workspace_bytes(100, 16, 16, 4) > 0
    # token slices of the weight-gradient kernel: 128 tokens per CTA up to 8192 rows, 192 above (workspace sizing)
    assert [lib.feta_linear_wgrad_slices(t) for t in (0, 1, 128, 129, 4736, 8192, 8193, 12032)] == \
        [1, 1, 1, 2, 37, 64, 43, 63]
    # argument validation happens before any CUDA call: a NULL pointer is rejected with FETA_EINVAL
    assert lib.feta_cheb_fwd(*([0] * 7), 0, 0, 0, 0, 0, 10, 1, 4, 16, 16, 8, 1, 0, 0, 0) == -1
    assert b"NULL" in lib.feta_last_error_string()


def test_sass_is_blackwell_native():
    """The shipped binary is sm_100a SASS and uses TMA bulk copies + packed fp32x2 FMAs."""
    so = os.path.join(ROOT, "feta_tmlr_b200", "libfeta_b200.so")
    elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    full = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    funcs = {}
    for chunk in full.split("Function : ")[1:]:
        funcs[chunk.split("\n", 1)[0].strip()] = chunk

    def sass_of(*needles):
        hits = [body for name, body in funcs.items() if all(n in name for n in needles)]
        assert hits, needles
        return "\n".join(hits)

    # warp-per-graph Chebyshev forward: TMA bulk copies (UBLKCP) + mbarrier waits (SYNCS) + packed fp32x2 FMAs,
    # and the F = 16 variant applies the filters on the tensor cores (3xTF32 HMMA.1688)
    sass = sass_of("cheb_fwd_warp_kernelILi16ELi2ELb1E")
    assert "UBLKCP" in sass and "FFMA2" in sass and "SYNCS" in sass and "HMMA.1688.F32.TF32" in sass
    sass = sass_of("cheb_fwd_warp_kernelILi8ELi2ELb0E")
    assert "UBLKCP" in sass and "FFMA2" in sass and "SYNCS" in sass
    # tcgen05 attention forward: tensor-core MMA (UTC*MMA), TMEM loads / stores
    sass = sass_of("attn_fwd_tc_kernelILi16E")
    assert "UTCHMMA" in sass and "LDTM" in sass and "STTM" in sass
    # weight gradients: the token-axis contraction runs on the tensor cores (3xTF32), 48 + 8 (db) MMAs per 32 tokens
    sass = sass_of("wgrad_partial_kernel")
    assert sass.count("HMMA.1688.F32.TF32") == 56 and "FFMA" not in sass


def test_weight_gradient_launches_are_dealt_over_the_side_streams(monkeypatch):
    """ops._wgrad_stream: round-robin over WGRAD_STREAMS side streams starting at the device's side stream 0; every
    stream that received work is a join target (ops.side_streams_in_use) until the end-of-pass join forgets it."""
    from feta_tmlr_b200 import ops
    made = {}
    monkeypatch.setattr(ops, "_side_stream", lambda device, index=0: made.setdefault((device, index), object()))
    monkeypatch.setattr(ops, "WGRAD_STREAMS", 3)
    monkeypatch.setattr(ops, "_WG_USED", {})
    monkeypatch.setattr(ops, "_WG_NEXT", {})
    dev = "cuda:0"
    assert ops.side_streams_in_use(dev) == []
    dealt = [ops._wgrad_stream(dev) for _ in range(7)]
    assert dealt[0] is made[(torch.device(dev), 0)]                    # the stream engine / models already know
    assert len({id(s) for s in dealt}) == 3 and dealt[:3] == dealt[3:6] and dealt[6] is dealt[0]
    assert ops.side_streams_in_use(dev) == dealt[:3]
    assert ops.side_streams_in_use(torch.device("cuda", 0)) == dealt[:3]          # one key per physical device
    assert ops.side_streams_in_use("cuda:1") == []
    monkeypatch.setattr(ops, "WGRAD_STREAMS", 1)
    monkeypatch.setattr(ops, "_WG_NEXT", {})
    assert all(ops._wgrad_stream(dev) is made[(torch.device(dev), 0)] for _ in range(3))


def test_no_cpu_fallback():
    import feta_tmlr_b200 as f
    m = f.ChebConvDynamic(4, 4, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(3, 4), torch.zeros((2, 0), dtype=torch.long), torch.zeros(2, 1, 4, 4), batch=torch.zeros(3))
    layer = f.DiffTransformerEncoderLayer(8, 2, 16, 0.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(torch.zeros(3, 1, 8), pe=None, degree=torch.ones(1, 3))
    with pytest.raises(NotImplementedError):
        f.ChebConvDynamic(4, 4, 2, aggr='mean')
    with pytest.raises(NotImplementedError):
        f.DiffTransformerEncoderGenGCN(8, 2, layer, 1, gnn_type='GENGCN')


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "feta_tmlr_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_state_dict_keys_match_reference_layout():
    import feta_tmlr_b200 as f
    m = f.DiffGraphTransformerGenGCNSBM(3, 2, 64, 4, 128, 0.0, 3)
    keys = set(m.state_dict())
    for k in ["embedding.weight", "encoder.layers.0.self_attn.in_proj_weight",
              "encoder.layers.2.self_attn.out_proj.weight", "encoder.layers.1.linear1.weight",
              "encoder.layers.1.norm2.bias", "encoder.spectral_gnns.bias", "encoder.gcn.weight",
              "encoder.gcn.bias", "encoder.linear.weight", "encoder.linear_cat.weight", "classifier.2.bias"]:
        assert k in keys, k
    assert m.state_dict()["encoder.gcn.weight"].shape == (1024, 1024)    # ncoef = K * dh^2 (models.py:133)
    lo = f.DiffTransformerEncoderGenGCN(64, 4, f.DiffTransformerEncoderLayer(64, 4, 128, 0.0), 2,
                                        learn_only_filter_order_coeff=True)
    assert lo.state_dict()["spectral_gnns.weight"].shape == (4, 16, 16)


_DDP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from feta_tmlr_b200 import ddp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
torch.manual_seed(rank)                       # replicas start different ...
model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
ddp.broadcast_parameters(model)               # ... and are made identical
bucket = ddp.FlatGradBucket(model.parameters())
g = torch.Generator().manual_seed(0)
X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
idx = list(ddp.shard_indices(8, rank, world))
bucket.zero()
((model(X[idx]) - Y[idx]) ** 2).sum().div(8).backward()
flat_local = bucket.flat.clone()
bucket.all_reduce_mean()
# single-process reference on the full batch with rank 0's weights
ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
ref.load_state_dict(model.state_dict())
((ref(X) - Y) ** 2).sum().div(8).backward()
ref_flat = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
assert torch.allclose(bucket.flat * world, ref_flat, atol=1e-6), (bucket.flat * world - ref_flat).abs().max()
assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
# epoch-end metric exchange: ragged shards come back concatenated in rank order on every rank
full_pred = torch.arange(11 * 3, dtype=torch.float32).view(11, 3)
full_lab = torch.arange(11)
mine = list(ddp.shard_indices(11, rank, world))
pred, lab = ddp.gather_predictions(full_pred[mine], full_lab[mine])
assert torch.equal(pred, full_pred) and torch.equal(lab, full_lab), (pred, lab)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_ddp_flat_bucket_gloo_world2(tmp_path):
    script = tmp_path / "ddp_worker.py"
    script.write_text(_DDP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_shard_indices_cover_everything():
    from feta_tmlr_b200 import ddp
    for n, w in [(10, 3), (8, 8), (5, 8), (1024, 4)]:
        parts = [list(ddp.shard_indices(n, r, w)) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--config", "MUTAG", "--batch", "4"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
