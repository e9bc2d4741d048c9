"""Row optimizer (8(f)-1): FusedRowOptimizer on the rows left by the fused step equals the torch
optimizer of the same name fed with the same sparse gradient."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("mode", ["rows", "apply"])
@pytest.mark.parametrize("learner,torch_cls,kw", [("sgd", torch.optim.SGD, {}), ("adagrad", torch.optim.Adagrad, {}),
                                                    ("sparse_adam", torch.optim.SparseAdam, {})])
def test_row_optimizer_matches_torch(learner, torch_cls, kw, mode):
    """'rows': gradient rows -> rsb200_rows_update; 'apply': the update fused into the scatter epilogue
    (RSB200_SINK_APPLY, no gradient rows).  Both equal the torch optimizer fed with the same sparse gradient."""
    from recstudio_b200 import retriever, rowopt
    U, N, d, B, n = 200, 3000, 64, 64, 50
    a = retriever.build_synthetic(U, N, d, n, fused_grad=mode, device=DEV, init_std=0.2, seed=1)
    b = retriever.build_synthetic(U, N, d, n, fused_grad="sparse", device=DEV, init_std=0.2, seed=1)
    assert torch.equal(a.item_encoder.weight, b.item_encoder.weight)
    opt_a = rowopt.FusedRowOptimizer(a, learner, lr=0.05)
    opt_b = torch_cls(b.parameters(), lr=0.05, **kw)
    gen = torch.Generator().manual_seed(0)
    for it in range(4):
        batch = {"user_id": torch.randint(1, U, (B,), generator=gen).to(DEV), "item_id": torch.randint(1, N, (B,), generator=gen).to(DEV),
                 "rating": torch.ones(B, device=DEV)}
        for m, opt in ((a, opt_a), (b, opt_b)):
            torch.manual_seed(100 + it)                       # identical negatives for both models
            opt.zero_grad()
            loss = m.training_step(dict(batch))
            loss.backward()
            opt.step()
        for wa, wb in ((a.item_encoder.weight, b.item_encoder.weight), (a.query_encoder.weight, b.query_encoder.weight)):
            scale = wb.abs().max().item()
            diff = (wa - wb).abs()
            if learner == "adagrad" or (learner == "sparse_adam" and mode == "apply"):
                # g / sqrt(sum g^2) (and Adam's m / sqrt(v)) is scale-free: an element whose gradient is a near-cancelled
                # sum (1e-9) amplifies the fused step's summation-order noise (the order of a row's entries comes from
                # atomics) to O(lr).  Almost all elements must agree to fp32 noise; the rest stay within a fraction of one
                # step.  ('rows' + sparse_adam shares the gradient kernel with the torch arm and is compared strictly.)
                assert (diff > 2e-6 * scale).float().mean().item() < 3e-3, (learner, it)
                assert diff.max().item() <= 0.1 * 0.05, (learner, it)
            else:
                assert diff.max().item() <= 2e-6 * scale, (learner, it)
    assert float(a.item_encoder.weight[0].abs().sum()) == 0.0     # padding row untouched


def test_row_optimizer_needs_rows():
    from recstudio_b200 import _lib, retriever, rowopt
    m = retriever.build_synthetic(20, 100, 32, 5, fused_grad="rows", device=DEV)
    with pytest.raises(_lib.Rsb200Error):
        rowopt.FusedRowOptimizer(m, "sgd").step()
    with pytest.raises(ValueError):
        rowopt.FusedRowOptimizer(m, "adam")


@pytest.mark.parametrize("loss,scorer", [("bpr", "ip"), ("ssm", "eu"), ("bpr", "eu")])
def test_apply_matches_rows(loss, scorer):
    """The fused epilogue (RSB200_SINK_APPLY) runs the very same per-element arithmetic (csrc/rowopt.cuh) on the same
    accumulated gradient as the two-kernel path; SGD is linear in the gradient, so the weights agree to fp32
    summation-order noise (d = 128: full rows; Euclid reads W[row] before updating it)."""
    from recstudio_b200 import retriever, rowopt
    U, N, d, B, n = 100, 2000, 128, 48, 70
    models = [retriever.build_synthetic(U, N, d, n, loss=loss, scorer=scorer, fused_grad=m, device=DEV, init_std=0.3, seed=3)
              for m in ("rows", "apply")]
    opts = [rowopt.FusedRowOptimizer(m, "sgd", lr=0.5) for m in models]
    gen = torch.Generator().manual_seed(1)
    for it in range(3):
        batch = {"user_id": torch.randint(1, U, (B,), generator=gen).to(DEV), "item_id": torch.randint(1, N, (B,), generator=gen).to(DEV),
                 "rating": torch.ones(B, device=DEV)}
        losses = []
        for m, opt in zip(models, opts):
            torch.manual_seed(7 + it)
            loss_t = m.training_step(dict(batch)); loss_t.backward(); opt.step()
            losses.append(loss_t.item())
        assert abs(losses[0] - losses[1]) <= 1e-6 * abs(losses[0])
        for name in ("item_encoder", "query_encoder"):
            wa, wb = getattr(models[0], name).weight, getattr(models[1], name).weight
            assert (wa - wb).abs().max().item() <= 1e-6 * wa.abs().max().item()
    assert float(models[1].item_encoder.weight[0].abs().sum()) == 0.0


def test_row_optimizer_is_a_torch_optimizer_with_state_dict_and_scheduler():
    """lr is read from param_groups at every step (torch lr schedulers drive it), the moments and the step
    count survive state_dict() / load_state_dict(), and the reference's _get_optimizers hook keeps the configured scheduler
    (returned under 'lr_scheduler' like recommender.py:431-441) and warns about semantics it cannot keep."""
    import warnings
    from recstudio_b200 import retriever, rowopt
    U, N, d, B, n = 50, 400, 32, 16, 8
    m = retriever.build_synthetic(U, N, d, n, fused_grad="rows", device=DEV, init_std=0.2, seed=1)
    m.config["train"].update({"learner": "adam", "learning_rate": 0.1, "weight_decay": 1e-4, "scheduler": "exponential"})
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        opts = m._get_optimizers()
    msgs = " ".join(str(x.message) for x in w)
    assert "SparseAdam" in msgs and "weight_decay" in msgs
    opt = opts[0]["optimizer"]
    assert isinstance(opt, rowopt.FusedRowOptimizer) and isinstance(opt, torch.optim.Optimizer)
    sched = opts[0]["lr_scheduler"]["scheduler"]
    assert isinstance(sched, torch.optim.lr_scheduler.ExponentialLR)
    batch = {"user_id": torch.arange(1, B + 1, device=DEV), "item_id": torch.arange(1, B + 1, device=DEV), "rating": torch.ones(B, device=DEV)}

    def one_step(model, o):
        torch.manual_seed(3)
        o.zero_grad(); model.training_step(dict(batch)).backward(); o.step()

    one_step(m, opt)
    sched.step()
    assert abs(opt.lr - 0.1 * 0.98) < 1e-12                        # the next step runs at the scheduled rate
    sd = opt.state_dict()
    assert sd["step_count"] == 1 and sd["state"]["item_encoder.weight"]["state1"].abs().sum().item() > 0
    # resume in a fresh model / optimizer: same weights + loaded state => identical second step
    m2 = retriever.build_synthetic(U, N, d, n, fused_grad="rows", device=DEV, init_std=0.2, seed=1)
    m2.load_state_dict(m.state_dict())
    opt2 = rowopt.FusedRowOptimizer(m2, "sparse_adam", lr=0.1)
    opt2.load_state_dict(sd)
    one_step(m, opt); one_step(m2, opt2)
    assert torch.equal(m.item_encoder.weight, m2.item_encoder.weight) and torch.equal(m.query_encoder.weight, m2.query_encoder.weight)
    with pytest.raises(ValueError):
        m.config["train"]["learner"] = "rmsprop"
        m._get_optimizers()


def test_graph_replay_behind_the_plugin_api():
    """fused_graph=True: training_step + backward replay two captured CUDA graphs (draw | COUNT | SCAN | FWD and SCATTER).
    Same negatives (the generator is advanced exactly like the eager path), same loss, bit-identical weights after
    FusedRowOptimizer steps; a changed batch size re-captures."""
    from recstudio_b200 import retriever, rowopt
    U, N, d, B, n = 300, 4000, 128, 64, 96
    models = [retriever.build_synthetic(U, N, d, n, loss="ssm", fused_grad="rows", fused_graph=g, device=DEV, init_std=0.2, seed=4)
              for g in (False, True)]
    opts = [rowopt.FusedRowOptimizer(m, "adagrad", lr=0.1) for m in models]
    gen = torch.Generator().manual_seed(2)
    for it, bsz in enumerate((B, B, B, 40, B)):
        batch = {"user_id": torch.randint(1, U, (bsz,), generator=gen).to(DEV), "item_id": torch.randint(1, N, (bsz,), generator=gen).to(DEV),
                 "rating": torch.ones(bsz, device=DEV)}
        losses, negs, ends = [], [], []
        for m, opt in zip(models, opts):
            torch.manual_seed(50 + it)
            opt.zero_grad()
            loss = m.training_step(dict(batch))
            assert type(loss.grad_fn).__name__.startswith("_GraphedStepFn" if m.fused_graph else "_FusedStepFn")
            loss.backward()
            opt.step()
            losses.append(loss.item()); negs.append(m.fused_last_neg_id().clone()); ends.append(torch.rand(2, device=DEV))
        assert torch.equal(negs[0], negs[1]) and torch.equal(ends[0], ends[1])
        assert losses[0] == losses[1]
        for name in ("item_encoder", "query_encoder"):
            assert torch.equal(getattr(models[0], name).weight, getattr(models[1], name).weight), (it, name)
