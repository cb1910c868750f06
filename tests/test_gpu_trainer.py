"""BASELINE config 1 on the GPU: the trainer's step sequence (finetune/finetuner.py:105-154) around a small random-init
two-tower CLIP — forward inside torch.autocast, projections of the last hidden states, SPARCLoss, every dict value
divided by gradient_accumulation_steps, GradScaler-scaled backward, then unscale_ + clip_grad_norm_ + step (here:
AdamSPD.amp_step, and the reference's own three calls) — for fp16 (the reference's AMP dtype) and bf16 autocast.
Step-1 loss and gradients are compared with the UNMODIFIED reference SPARCLoss (oracle/_ref) run eagerly on the same
GPU on the same embeddings."""
import copy
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny_clip():
    from transformers import CLIPConfig, CLIPModel
    cfg = CLIPConfig(
        text_config=dict(vocab_size=1000, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         max_position_embeddings=77, pad_token_id=1, bos_token_id=998, eos_token_id=999),
        vision_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=64,
                           patch_size=16),
        projection_dim=128)
    torch.manual_seed(0)
    return CLIPModel(cfg)


def _batch(B, dev):
    g = torch.Generator().manual_seed(1)
    pixel = torch.randn(B, 3, 64, 64, generator=g)
    ids = torch.randint(2, 990, (B, 77), generator=g)          # no pad token (id 1): all-True mask, as in real runs
    ids[:, 0] = 998; ids[:, -1] = 999
    return pixel.to(dev), ids.to(dev)


def _embeddings(model, pixel, ids):
    out = model(pixel_values=pixel, input_ids=ids)
    v = model.visual_projection(out.vision_model_output.last_hidden_state)     # finetuner.py:125-126  [B, 1 + P, D]
    l = model.text_projection(out.text_model_output.last_hidden_state)         # finetuner.py:127-128  [B, T, D]
    return v, l


@pytest.mark.parametrize("amp_dtype", [torch.float16, torch.bfloat16])
def test_trainer_sequence_three_steps(amp_dtype):
    pytest.importorskip("transformers")
    from clip_finegrained_alignment_b200 import AdamSPD, SPARCLoss
    from oracle import make_ref
    dev = torch.device("cuda")
    B, accum, max_norm = 8, 2, 1.0
    model = _tiny_clip().to(dev)
    pixel, ids = _batch(B, dev)
    mask = torch.ne(ids, model.config.text_config.pad_token_id).bool()
    P = (64 // 16) ** 2 + 1
    cfg = types.SimpleNamespace(similarity_threshold=1.0 / P, global_loss_weight=1.0, local_loss_weight=1.0,
                                inverse_temperature=1.0)
    crit = SPARCLoss(cfg).to(dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = AdamSPD([{"params": params, "pre": copy.deepcopy(params)}], lr=1e-3, weight_decay=0.1)     # finetuner.py:81-101
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)

    # ---- step 1, checked against the unmodified reference loss on the same embeddings
    with torch.autocast("cuda", dtype=amp_dtype):
        v, l = _embeddings(model, pixel, ids)
        assert v.dtype == amp_dtype and l.dtype == amp_dtype
        losses = crit(v, l, mask)
        losses = {k: x / accum for k, x in losses.items()}                      # finetuner.py:145
    scaler.scale(losses["total_loss"]).backward()
    ref_mods = make_ref.import_ref()
    if ref_mods is not None:
        vr = v.detach().float().requires_grad_(True)
        lr_ = l.detach().float().requires_grad_(True)
        rl = ref_mods[0].SPARCLoss(cfg).to(dev)(vr, lr_, mask)
        for k in rl:
            assert abs(float(losses[k]) * accum - float(rl[k])) <= 2e-4 * max(1.0, abs(float(rl[k]))), (k, float(losses[k]) * accum, float(rl[k]))
        # gradient w.r.t. the embeddings of the scaled, accumulated loss (what our backward received): compare through
        # the gradient that reached the projection weights
        (rl["total_loss"] / accum * scaler.get_scale()).backward()
        gv_ref = torch.einsum("bpd,bph->dh", vr.grad, model.vision_model(pixel).last_hidden_state.float().detach())
        gv = model.visual_projection.weight.grad.float()
        rel = float((gv - gv_ref).norm() / gv_ref.norm())
        assert rel <= (2e-2 if amp_dtype == torch.float16 else 4e-2), rel           # autocast rounds the projection GEMMs
    first = float(losses["total_loss"]) * accum
    total_norm = opt.amp_step(scaler, max_norm)                                 # unscale_ + clip_grad_norm_ + scaler.step
    scaler.update()
    opt.zero_grad()
    assert torch.isfinite(total_norm)

    # ---- two more steps: the loss goes down
    last = first
    for _ in range(2):
        with torch.autocast("cuda", dtype=amp_dtype):
            v, l = _embeddings(model, pixel, ids)
            losses = {k: x / accum for k, x in crit(v, l, mask).items()}
        scaler.scale(losses["total_loss"]).backward()
        opt.amp_step(scaler, max_norm)
        scaler.update()
        opt.zero_grad()
        last = float(losses["total_loss"]) * accum
    assert last < first, (first, last)
    for p in params:
        assert torch.isfinite(p).all()


def test_reference_three_call_sequence_equals_amp_step():
    """scaler.unscale_(opt); clip_grad_norm_; scaler.step(opt) (finetuner.py:150-152) on AdamSPD == AdamSPD.amp_step."""
    pytest.importorskip("transformers")
    from clip_finegrained_alignment_b200 import AdamSPD, SPARCLoss
    dev = torch.device("cuda")
    B = 4
    outs = []
    for mode in ("calls", "amp_step"):
        model = _tiny_clip().to(dev)
        pixel, ids = _batch(B, dev)
        mask = torch.ne(ids, 1).bool()
        cfg = types.SimpleNamespace(similarity_threshold=1.0 / 17, global_loss_weight=1.0, local_loss_weight=1.0,
                                    inverse_temperature=1.0)
        crit = SPARCLoss(cfg)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = AdamSPD([{"params": params, "pre": copy.deepcopy(params)}], lr=1e-3, weight_decay=0.1)
        scaler = torch.amp.GradScaler("cuda", init_scale=256.0)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            v, l = _embeddings(model, pixel, ids)
            loss = crit(v, l, mask)["total_loss"]
        scaler.scale(loss).backward()
        if mode == "calls":
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            scaler.step(opt)
        else:
            opt.amp_step(scaler, 1.0)
        scaler.update()
        outs.append([p.detach().clone() for p in params])
    for a, b in zip(*outs):
        assert float((a - b).abs().max()) <= 1e-6
