"""CPU-only checks (no GPU, no compute calls): the C-ABI library loads and exports every
declared symbol, the ctypes mirror matches the header, host-side logic (Philox offset
arithmetic, workspace sizing, argument validation) and the plugin class hierarchy."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import REPO
from oracle import philox

REF = "/root/reference"


@pytest.fixture(scope="module")
def L():
    from recstudio_b200 import _lib, build
    build.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(L):
    from recstudio_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 20
    missing = [s for s in names if not hasattr(L, s)]
    assert not missing, missing
    assert L.rsb200_version() == 200
    assert L.rsb200_sizeof_pair_args() == C.sizeof(_lib.PairArgs)


def test_sass_is_sm100a_only():
    from recstudio_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {line.split(".")[-2] for line in out.split() if line.endswith(".cubin")}
    assert archs == {"sm_100a"}, archs


def test_philox_counter_offset_matches_oracle(L):
    for numel in (1, 255, 256, 257, 1184 * 256 * 4, 1184 * 256 * 4 + 1, 8192 * 1024, 99_999_999):
        for sm, mt in ((148, 2048), (132, 2048), (108, 1536)):
            assert L.rsb200_philox_counter_offset(numel, sm, mt) == philox.torch_cuda_counter_offset(numel, sm, mt)
    assert L.rsb200_philox_counter_offset(0, 148, 2048) == 0


def test_workspace_sizes_and_argument_errors(L):
    from recstudio_b200 import _lib
    sz = _lib.PairSizes()
    assert L.rsb200_pair_workspace_sizes(10_000_001, 1_000_001, 8192, 1024, 128, C.byref(sz)) == 0
    assert sz.off_item == 10_000_002 and sz.ent_item == 8192 * 1025 and sz.cap_item == 8192 * 1025
    assert sz.cap_user == 8192 and sz.q_buf == 8192 * 128 and sz.scan_tmp >= 10_000_001 // 4096 + 1
    assert L.rsb200_pair_workspace_sizes(100, 100, 4, 4, 6, C.byref(sz)) == -1          # d % 4 != 0
    assert b"bad problem shape" in L.rsb200_last_error()
    assert L.rsb200_pair_workspace_sizes(100, 100, 1 << 22, 1 << 10, 8, C.byref(sz)) == -2   # B*(n+1) >= 2^31
    # null / inconsistent argument blocks are rejected before any CUDA call
    a = _lib.PairArgs()
    assert L.rsb200_pair_step(C.byref(a), 15, None) == -1
    assert L.rsb200_pair_step(None, 15, None) == -1
    # sampler argument validation (no device needed: checks precede the launch)
    assert L.rsb200_sample_uniform(1, 3, 100, 4, 4, 148, 2048, None, None, None) == -1       # offset % 4
    assert L.rsb200_sample_uniform(1, 0, 1, 4, 4, 148, 2048, None, None, None) == -1         # no items
    assert L.rsb200_sample_uniform(1, 0, (1 << 28) + 2, 4, 4, 148, 2048, None, None, None) == -2   # 64-bit draw path
    assert L.rsb200_sample_uniform(1, 0, 100, 1 << 20, 1 << 9, 148, 2048, None, None, None) == -2  # numel*8 >= 2^31


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA (this container has no GPU)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from recstudio_b200 import _lib, fused, plugins, sampling
    with pytest.raises(_lib.Rsb200Error):
        fused.PairWorkspace(10, 10, 2, 2, 8, "cpu")
    with pytest.raises(_lib.Rsb200Error):
        sampling.uniform_draw(10, 2, 2, "cpu")
    emb = plugins.FusedEmbedding(10, 8)
    with pytest.raises(_lib.Rsb200Error):
        emb(torch.tensor([1, 2]))
    with pytest.raises(_lib.Rsb200Error):
        plugins.FusedBPRLoss()(None, torch.zeros(2), None, torch.zeros(2, 3), None)
    with pytest.raises(_lib.Rsb200Error):
        plugins.FusedInnerProductScorer()(torch.zeros(2, 8), torch.zeros(2, 8))
    with pytest.raises(_lib.Rsb200Error):
        plugins.FusedUniformSampler(10)(torch.zeros(2, 8), 3)
    with pytest.raises(_lib.Rsb200Error):
        plugins.FusedEmbedding(10, 6)            # rows must be 16-byte multiples


def test_product_never_imports_the_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, "recstudio_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                if "import oracle" in src or "from oracle" in src:
                    bad.append(f)
    assert not bad, bad


def test_popular_sampler_tables_match_reference_golden():
    from conftest import load_golden
    from recstudio_b200 import plugins
    g = load_golden("popular")
    for tag in ("small", "big"):
        for mode in (0, 1, 2):
            s = plugins.FusedPopularSampler(g[f"{tag}_count"], mode=mode)
            np.testing.assert_array_equal(s.pop_prob.numpy(), g[f"{tag}_m{mode}_prob"])
            np.testing.assert_array_equal(s.table.numpy(), g[f"{tag}_m{mode}_table"])
            assert s.num_items == g[f"{tag}_count"].shape[0] - 1


def test_fused_retriever_is_the_reference_retriever():
    """recstudio_b200.iface binds the REAL reference classes (installed under baseline/_ref by
    __graft_entry__.build()): every plugin passes the reference's isinstance gates, BaseRetriever's kwargs
    constructor accepts them unchanged, and the rank metrics / _test_step / forward / _sample the retriever
    runs are the reference's own functions (no mirrors in this package)."""
    code = r"""
import sys, os, tempfile
sys.path.insert(0, %r)
os.chdir(tempfile.mkdtemp())
import warnings; warnings.filterwarnings('ignore')
import torch, pytest
from recstudio_b200 import iface, plugins, retriever
import recstudio
from recstudio.model.basemodel import BaseRetriever
from recstudio.ann.sampler import Sampler
from recstudio.model import loss_func, scorer, init
assert issubclass(plugins.FusedUniformSampler, Sampler) and issubclass(plugins.FusedPopularSampler, Sampler)
assert issubclass(plugins.FusedBPRLoss, loss_func.PairwiseLoss) and issubclass(plugins.FusedSampledSoftmaxLoss, loss_func.PairwiseLoss)
assert issubclass(plugins.FusedSoftmaxLoss, loss_func.FullScoreLoss)
assert issubclass(plugins.FusedInnerProductScorer, scorer.InnerProductScorer)
assert issubclass(plugins.FusedEuclideanScorer, scorer.EuclideanScorer)
assert issubclass(plugins.FusedEmbedding, torch.nn.Embedding)
assert issubclass(retriever.FusedRetriever, BaseRetriever)
# the reference's own code, not copies: these attributes resolve to functions defined in recstudio
for name in ('forward', '_sample', '_test_step', 'validation_step', 'test_step', '_update_item_vector', '_to_device', 'fit'):
    assert getattr(retriever.FusedRetriever, name).__module__.startswith('recstudio.'), name
assert not os.path.exists(os.path.join(os.path.dirname(retriever.__file__), 'rank_metrics.py'))
from recstudio.utils import get_model
conf = get_model('BPR')[1]; conf['train']['gpu'] = None; conf['train']['negative_count'] = 4; conf['model']['embed_dim'] = 8
m = retriever.FusedRetriever(conf, fused_grad='sparse', item_encoder=plugins.FusedEmbedding(30, 8), query_encoder=plugins.FusedEmbedding(12, 8),
        scorer=plugins.FusedInnerProductScorer(), sampler=plugins.FusedUniformSampler(30), loss=plugins.FusedBPRLoss())
# the reference's constructor asserts (baseretriever.py:17-41, recommender.py:48-54)
for bad in (dict(sampler=object()), dict(loss=torch.nn.Identity()), dict(item_encoder=3)):
    with pytest.raises(AssertionError):
        retriever.FusedRetriever(conf, **bad)
m.fiid = 'item_id'
with pytest.raises(NotImplementedError):                      # the reference's message for an unknown method
    m.sampling({'user_id': torch.tensor([1]), 'item_id': torch.tensor([1])}, 3, method='is')
with pytest.raises(AssertionError):                           # baseretriever.py:266-270
    m.sampling({'user_id': torch.tensor([1]), 'item_id': torch.tensor([1])}, [2, 3], method='dns')
# reference parameter init dispatches on isinstance(nn.Embedding) and re-zeroes the padding row (init.py:5-9)
m.item_encoder.apply(init.xavier_normal_initialization)
assert float(m.item_encoder.weight[0].abs().sum()) == 0.0 and float(m.item_encoder.weight[1:].abs().sum()) > 0
# state_dict exposes the tables under the reference's key names (checkpoint compatibility)
assert 'item_encoder.weight' in m.state_dict() and 'query_encoder.weight' in m.state_dict()
# hooks used when no kwargs are given
class D: num_items = 30; num_users = 12
m2 = retriever.FusedBPR(conf)
assert isinstance(m2._get_item_encoder(D), plugins.FusedEmbedding) and isinstance(m2._get_sampler(D), plugins.FusedUniformSampler)
assert isinstance(m2._get_loss_func(), plugins.FusedBPRLoss) and isinstance(m2.score_func, plugins.FusedInnerProductScorer)
print('OK', os.path.dirname(recstudio.__file__))
""" % (REPO,)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_reference_install_is_unmodified():
    """baseline/_ref is the pip --target copy of /root/reference: byte-identical sources (checked where the read-only
    checkout exists, i.e. in the authoring container)."""
    ref_src, inst = os.path.join(REF, "recstudio"), os.path.join(REPO, "baseline", "_ref", "recstudio")
    if not os.path.isdir(ref_src):
        pytest.skip("needs the read-only reference checkout")
    assert os.path.isdir(inst), "run __graft_entry__.build() first"
    import filecmp
    checked = 0
    for root, _, files in os.walk(inst):
        for f in files:
            if f.endswith((".py", ".yaml")):
                a, b = os.path.join(root, f), os.path.join(ref_src, os.path.relpath(os.path.join(root, f), inst))
                assert filecmp.cmp(a, b, shallow=False), a
                checked += 1
    assert checked > 100


def test_c_abi_from_plain_c(tmp_path):
    """include/rsb200.h is a C header and librsb200.so a plain C-ABI library: a C99 program with no Python / torch in
    the process links it and exercises the size queries, argument validation and error strings."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    from recstudio_b200 import build
    lib = build.build()
    exe = str(tmp_path / "abi_smoke")
    src = os.path.join(REPO, "tests", "c_abi", "abi_smoke.c")
    libdir = os.path.dirname(lib)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), src, "-o", exe,
                    "-L", libdir, "-lrsb200", "-Wl,-rpath," + libdir], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "abi ok" in out.stdout, out.stdout + out.stderr


def test_sequence_pooling_matches_reference_golden():
    """attention.pool_sequence == the reference's SeqPoolingLayer for all seven pooling types (tests/golden/pooling.npz);
    device-agnostic torch plumbing around the attention kernels, so it is checked here on the CPU."""
    import numpy as np
    from conftest import load_golden
    from recstudio_b200 import attention
    g = load_golden("pooling")
    x, seqlen, mask = torch.from_numpy(g["x"]), torch.from_numpy(g["seqlen"]), torch.from_numpy(g["mask_token"])
    for ptype in ("origin", "mask", "concat", "sum", "mean", "last"):
        got = attention.pool_sequence(x, seqlen, ptype, mask if ptype == "mask" else None)
        np.testing.assert_array_equal(got.numpy(), g[ptype])
    mx = attention.pool_sequence(x, seqlen, "max")
    np.testing.assert_array_equal(mx.values.numpy(), g["max_values"])
    np.testing.assert_array_equal(mx.indices.numpy(), g["max_indices"])
    with pytest.raises(ValueError):
        attention.pool_sequence(x, seqlen, "first")


def test_every_declared_function_has_a_typed_binding(L):
    """The ctypes layer declares argument types for every entry point of include/rsb200.h, and the number of arguments in
    each binding equals the number of parameters in the C declaration (a drifted binding would corrupt the call frame)."""
    import re
    from recstudio_b200 import _lib
    src = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER).read(), flags=re.S)
    decls = dict((m.group(1), m.group(2)) for m in re.finditer(r"\b(rsb200_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert set(decls) == set(_lib.declared_symbols())
    for name, params in sorted(decls.items()):
        fn = getattr(L, name)
        params = params.strip()
        n_params = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert fn.argtypes is not None or n_params == 0, "no argtypes for %s" % name
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n_params, "%s: binding has %d arguments, header declares %d" % (name, len(fn.argtypes), n_params)


def test_popular_slice_tables_match_reference_golden():
    """sharded.PopularSlice.tables (what every rank builds on the CPU before slicing) == PopularSamplerModel's buffers as the
    unmodified reference produced them (tests/golden/popular.npz, sampler.py:225-241)."""
    from conftest import load_golden
    from recstudio_b200 import sharded
    g = load_golden("popular")
    for tag in ("small", "big"):
        for mode in (0, 1, 2):
            table, prob = sharded.PopularSlice.tables(g[f"{tag}_count"], mode)
            np.testing.assert_array_equal(prob.numpy(), g[f"{tag}_m{mode}_prob"])
            np.testing.assert_array_equal(table.numpy(), g[f"{tag}_m{mode}_table"])


def test_uniform_owner_range_is_exact(L):
    """rsb200_uniform_owner_range (host arithmetic of the owner-side UniformSampler regeneration, sampler.py:86-111 =
    randint(1, N)): for every 32-bit word v,  lo <= low64(magic * v) <= hi  <=>  row0 <= v mod (N - 1) + 1 < row0 + local_rows,
    and mulhi64(low64(magic * v), N - 1) == v mod (N - 1).  Checked with Python integers on random and boundary words, every
    owner of several world sizes, table sizes from 2 rows to 2^32."""
    rng = np.random.default_rng(5)
    sizes = [2, 3, 4, 10, 1000, 1575, 400_003, 1_000_001, 10_000_001, 100_000_001, 2 ** 31 - 5, 2 ** 31 + 1, 2 ** 32 - 1, 2 ** 32,
             2 ** 20 + 1, 2 ** 24 + 1]
    sizes += [int(x) for x in rng.integers(2, 2 ** 32, size=24)]
    mask = (1 << 64) - 1
    for N in sizes:
        d = N - 1
        for world in (1, 2, 3, 8):
            per = -(-N // world)
            for r in range(world):
                row0 = r * per
                local = max(0, min(per, N - row0))
                if row0 > N:
                    continue
                mg, lo, hi = C.c_uint64(), C.c_uint64(), C.c_uint64()
                assert L.rsb200_uniform_owner_range(N, row0, local, C.byref(mg), C.byref(lo), C.byref(hi)) == 0
                M = mg.value
                assert M == -(-(1 << 64) // d) or (d & (d - 1)) == 0 and M == ((1 << 64) // d) & mask
                words = [int(x) for x in rng.integers(0, 2 ** 32, size=48)] + [0, 1, 2 ** 32 - 1, d % 2 ** 32, (d - 1) % 2 ** 32]
                for b in (row0 - 2, row0 - 1, row0, row0 + local - 2, row0 + local - 1, row0 + local):   # ids around the block's ends
                    if 0 <= b < d:
                        words += [b, (b + d) % 2 ** 32] if b + d < 2 ** 32 else [b]
                for v in words:
                    low = (M * v) & mask
                    assert (low * d) >> 64 == v % d, (N, v)
                    owned = row0 <= v % d + 1 < row0 + local
                    assert owned == (lo.value <= low <= hi.value), (N, world, r, v)
    mg, lo, hi = C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert L.rsb200_uniform_owner_range(2 ** 32 + 2, 0, 10, C.byref(mg), C.byref(lo), C.byref(hi)) == -1   # ids beyond 32-bit words
    assert L.rsb200_uniform_owner_range(10, 0, 10, None, C.byref(lo), C.byref(hi)) == -1


def test_device_batch_loader_yields_the_reference_loaders_batches():
    """loader.DeviceBatchLoader (here on the CPU device: it is plain torch indexing) delivers, batch for batch and
    bit for bit, what the REFERENCE's own `train_loader()` delivers for the same global seed -- ml-100k as bundled with the
    reference, B = 512, two epochs incl. the ragged last batch (dataset.py:1083-1123,1687-1734: DataSampler's per-epoch CPU
    randperm seeded from the global generator, after the DataLoader iterator drew its base seed)."""
    code = r"""
import sys, os, tempfile
sys.path.insert(0, %r)
os.chdir(tempfile.mkdtemp())
import warnings; warnings.filterwarnings('ignore')
import torch
from recstudio_b200 import iface, loader
from recstudio.data.dataset import TripletDataset
from recstudio.utils import get_model
conf = get_model('BPR')[1]
data_conf = {'user_feat_name': None}; data_conf.update(conf['data'])
trn = TripletDataset(name='ml-100k', config=data_conf).build(**conf['data'])[0]
for bs, drop_last in ((512, False), (1000, True)):
    torch.manual_seed(2022)
    ref = [{k: v.clone() for k, v in b.items() if isinstance(v, torch.Tensor)}
           for _ in range(2) for b in trn.train_loader(batch_size=bs, shuffle=True, num_workers=0, drop_last=drop_last)]
    torch.manual_seed(2022)
    dl = loader.DeviceBatchLoader.from_dataset(trn, bs, 'cpu', drop_last=drop_last)
    mine = [b for _ in range(2) for b in dl]
    assert len(ref) == len(mine) == 2 * len(dl) and len(ref) > 100 // (bs // 512), (len(ref), len(mine), len(dl))
    for a, b in zip(ref, mine):
        assert set(a) == set(b) == {'user_id', 'item_id', 'rating', 'timestamp'}
        for k in a:
            assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    if not drop_last:
        assert ref[len(dl) - 1]['user_id'].shape[0] == len(trn) %% bs           # the ragged tail batch
print('OK', len(trn))
""" % (REPO,)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
