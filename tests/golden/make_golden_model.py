"""Goldens for the MODEL-LEVEL code, produced by running the reference's own unmodified Python:

    playground/detection.3d/waymo/conquer/VoxelDETR.../{net,voxel_detr,transformer,heads,losses}.py + modules/*
    playground/detection.3d/waymo/conquer/ConQueR.../{net,voxel_detr,transformer,heads,losses,cdn}.py + modules/*
    playground/detection.3d/waymo/center_point/centerpoint.../{net,voxelnet,center_head,centernet_loss,...}.py
    efg/modeling/backbones/{sparse_net,fpn,configurable_rpn}.py, efg/modeling/readers/voxel_reader.py, ...

through `net.build_model(None, config)` and `model(batched_inputs)` (the plugin surface of SURVEY.md §8b), on the CPU of
the build container: spconv -> oracle/spconv_cpu.py, BoxAttnFunction -> the reference's own torch twin
(efg/operators/ms_deform_attn.py:55-76); see tests/golden/ref_env.py for the import environment.

Stored per model (tests/golden/model_<kind>.pt): the seeded input scenes, the training losses, the gradients of a few
parameters and gradient norms of all, the reference's state_dict key -> shape table, for ConQueR the random draws of
prepare_for_cdn, and the eval-mode detections.  Weights are NOT stored: tests/golden/model_cases.py:fill_state_dict
regenerates them from the parameter names.

Usage: python tests/golden/make_golden_model.py [voxel_detr conquer centerpoint]
"""
import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import model_cases as mc  # noqa: E402
import ref_env  # noqa: E402

EXP = {"voxel_detr": ref_env.VD_DIR, "conquer": ref_env.CQ_DIR, "centerpoint": ref_env.CP_DIR}


class _RecordRandom:
    """Record what torch.rand_like / torch.randint_like return (clones), in call order."""

    def __enter__(self):
        self.draws = []
        self._rand_like, self._randint_like = torch.rand_like, torch.randint_like

        def rand_like(*a, **k):
            out = self._rand_like(*a, **k)
            self.draws.append(("rand_like", out.clone()))
            return out

        def randint_like(*a, **k):
            out = self._randint_like(*a, **k)
            self.draws.append(("randint_like", out.clone()))
            return out

        torch.rand_like, torch.randint_like = rand_like, randint_like
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.randint_like = self._rand_like, self._randint_like


def run_reference(kind):
    from oracle import spconv_cpu

    cfg = mc.make_config(kind)
    scenes = mc.make_scenes(kind)
    with ref_env.playground(EXP[kind], spconv_module=spconv_cpu):
        from net import build_model  # the reference's plugin entry point (cli/main.py:120,144)

        torch.manual_seed(0)
        model = build_model(None, ref_env.to_cfg(copy.deepcopy(dict(cfg))))
        keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(mc.fill_state_dict(model.state_dict()))
        model.train()
        torch.manual_seed(1)
        with ref_env.cuda_calls_stay_on_cpu(), _RecordRandom() as rec:
            losses = model(mc.make_batch(scenes, cfg.dataset))
        total = sum(v for k, v in losses.items() if "loss" in k and v.requires_grad)
        total.backward()
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        out = {
            "kind": kind,
            "scenes": scenes,
            "state_dict_shapes": keys,
            "losses": {k: float(v.detach()) for k, v in losses.items()},
            "total": float(total.detach()),
            "grads": {k: grads[k] for k in mc.GRAD_KEYS[kind] if k in grads},
            "grad_norms": {k: float(g.norm()) for k, g in grads.items()},
            "no_grad_params": sorted(n for n, p in model.named_parameters() if p.grad is None),
        }
        missing = [k for k in mc.GRAD_KEYS[kind] if k not in grads]
        assert not missing, missing
        if kind == "conquer":
            d = rec.draws
            assert [n for n, _ in d[:4]] == ["rand_like", "randint_like", "randint_like", "rand_like"], [n for n, _ in d]
            out["cdn_draws"] = {"p_label": d[0][1], "new_label_chosen": d[1][1], "rand_sign": d[2][1], "rand_part": d[3][1]}
        if kind in ("voxel_detr", "conquer"):
            # eval on the pristine weights (the training forward above moved the BN running statistics and the EMA decoder)
            model.load_state_dict(mc.fill_state_dict(model.state_dict()))
            model.eval()
            with torch.no_grad(), ref_env.cuda_calls_stay_on_cpu():
                res = model(mc.make_batch(scenes[:1], cfg.dataset))
            out["eval"] = [{k: v.clone() for k, v in r.items()} for r in res]
    return out


if __name__ == "__main__":
    for kind in (sys.argv[1:] or ["voxel_detr", "conquer", "centerpoint"]):
        g = run_reference(kind)
        path = os.path.join(HERE, "model_%s.pt" % kind)
        torch.save(g, path)
        print(kind, "losses:", {k: round(v, 5) for k, v in g["losses"].items()})
        print(kind, "%d keys, %d params with grad, file %.1f KB" % (len(g["state_dict_shapes"]), len(g["grad_norms"]),
                                                                   os.path.getsize(path) / 1e3))
