"""Time one PPO minibatch update (take + forward + loss + backward + clip/Adam [+ re-pack]) at ppo.py's
minibatch size, library-GEMM back end vs the tcgen05 back end (wide / narrow tiles).  B200 only.

    python scripts/exp_update.py [--mbs 1024] [--iters 200]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbs", type=int, default=1024)
    ap.add_argument("--total", type=int, default=8192 * 32)
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    from brl_b200 import ops
    from brl_b200.models import LAYERS, init_params
    from brl_b200.optim import AdamWithClip, flatten_params
    from brl_b200.update import _LossHead, _forward_autograd
    dev = "cuda:0"
    B, total = args.mbs, args.total
    g = torch.Generator().manual_seed(0)
    obs = (torch.rand((total, 480), generator=g) < 0.05).to(torch.bfloat16).to(dev)
    mask = (torch.rand((total, 38), generator=g) < 0.5)
    mask[:, 0] = True
    mask = mask.to(torch.uint8).to(dev)
    action = torch.zeros(total, dtype=torch.int32, device=dev)
    old_lp = (-torch.rand(total, generator=g) * 3).to(dev)
    old_v = (torch.randn(total, generator=g) * 0.3).to(dev)
    adv = torch.randn(total, generator=g).to(dev)
    tgt = (torch.randn(total, generator=g) * 0.3).to(dev)
    perm = torch.randperm(total, generator=g).to(torch.int32).to(dev)
    params = init_params(1, dev)
    cfg = dict(clip_eps=0.2, ent_coef=0.01, vf_coef=0.5, illegal_l2_coef=0.0, value_clipping=True, reward_scaling=False,
               masked_policy=True)
    opt = AdamWithClip(1e-4, eps=1e-5, max_grad_norm=0.5)
    nmb = total // B
    out = {"mbs": B, "iters": args.iters}

    def timed(fn):
        for i in range(10):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    # --- tensor-core back end
    flat_p, _ = flatten_params(params)
    state = opt.init(params)
    blob = ops.mlp_pack_train(flat_p)
    scratch = ops.mlp_train_scratch(B, dev)
    flat_g = torch.empty_like(flat_p)
    stats = torch.zeros(8, dtype=torch.float32, device=dev)
    acc = ops.ppo_scratch(dev)
    for tune, name in ((0, "tc_fused"), (16, "tc_fused_narrow_bwd"), (4, "tc_fused_wide_fwd")):
        st = [state]

        def step(i, tune=tune):
            mb = i % nmb
            ops.ppo_grad(obs, blob, scratch, perm[mb * B:(mb + 1) * B], mask, action, old_lp, old_v, adv, tgt, flat_g, stats, acc,
                         tune=tune, **cfg)
            st[0] = opt.update_mlp_(flat_p, flat_g, st[0], acc[14:15], blob)

        def grad_only(i, tune=tune):
            mb = i % nmb
            ops.ppo_grad(obs, blob, scratch, perm[mb * B:(mb + 1) * B], mask, action, old_lp, old_v, adv, tgt, flat_g, stats, acc,
                         tune=tune, **cfg)

        out[name + "_ms_per_update"] = timed(step)
        out[name + "_ms_grad_only"] = timed(grad_only)
        # the same update replayed from a CUDA graph (launch overhead removed)
        gph = torch.cuda.CUDAGraph()
        idx_static = perm[:B].clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ops.ppo_grad(obs, blob, scratch, idx_static, mask, action, old_lp, old_v, adv, tgt, flat_g, stats, acc, tune=tune, **cfg)
            torch.cuda.synchronize()
            with torch.cuda.graph(gph, stream=s):
                ops.ppo_grad(obs, blob, scratch, idx_static, mask, action, old_lp, old_v, adv, tgt, flat_g, stats, acc, tune=tune, **cfg)
        torch.cuda.current_stream().wait_stream(s)
        out[name + "_ms_grad_only_graph"] = timed(lambda i: gph.replay())
    # --- library-GEMM back end (cuBLAS fp32 through torch autograd)
    flat_p2, new_params = flatten_params(params)
    leaves = {name: {k: new_params[name][k].detach().requires_grad_() for k in ("w", "b")} for name in LAYERS}
    flat_g2 = torch.zeros_like(flat_p2)
    off = 0
    for name in LAYERS:
        for k in ("w", "b"):
            t = leaves[name][k]
            t.grad = flat_g2[off: off + t.numel()].view(t.shape)
            off += t.numel()
    call = dict(mask=mask, action=action, old_log_prob=old_lp, old_value=old_v, adv=adv, targets=tgt, cfg=cfg, scratch=acc, stats=stats)
    x_mb = torch.empty((B, 480), dtype=obs.dtype, device=dev)
    st2 = [opt.init(params)]
    torch.backends.cuda.matmul.allow_tf32 = False

    def step_lib(i):
        mb = i % nmb
        index = perm[mb * B:(mb + 1) * B]
        ops.gather_rows(obs, index, x_mb)
        logits, value = _forward_autograd(leaves, x_mb.to(torch.float32), torch.relu)
        call["index"] = index
        loss = _LossHead.apply(logits.contiguous(), value.contiguous(), call)
        flat_g2.zero_()
        loss.backward()
        st2[0] = opt.update_(flat_p2, flat_g2, st2[0])

    out["library_fp32_ms_per_update"] = timed(step_lib)
    flop = 3 * 7354368 * B
    for k in list(out):
        if k.endswith("_ms_per_update") or "grad_only" in k:
            out[k.replace("_ms_", "_model_TFLOPs_")] = flop / (out[k] * 1e-3) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
