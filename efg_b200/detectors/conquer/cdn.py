"""Contrastive denoising queries of ConQueR (CQ/cdn.py:5-139).

For every ground-truth box, ``dn_number`` groups of one positive (noise < 1 half-size) and one
negative (noise in [1, 2) half-sizes) noised copy are prepended to the decoder queries; an attention
mask keeps groups, and the matching queries, from seeing each other.  Per scene the layout is
``position = repeat * single_pad + local_gt_index`` with repeat 2g = positives, 2g+1 = negatives of group g.

The random draws can be supplied (``noise`` dict with keys p_label, new_label, rand_sign, rand_part) so
that two implementations run on identical noise; otherwise they are drawn on
the targets' device.  Device-agnostic (the reference hard-codes ``.cuda()``).
"""
import torch
from torch.nn import functional as F


def draw_noise(num_rows, num_classes, device, generator=None):
    """The four random tensors prepare_for_cdn consumes, for num_rows = 2 * dn_number * num_gt rows."""
    return {
        "p_label": torch.rand(num_rows, device=device, generator=generator),
        "new_label": torch.randint(0, num_classes, (num_rows,), device=device, generator=generator),
        "rand_sign": torch.randint(0, 2, (num_rows, 7), device=device, generator=generator).float() * 2.0 - 1.0,
        "rand_part": torch.rand(num_rows, 7, device=device, generator=generator),
    }


def prepare_for_cdn(targets, dn_number, label_noise_ratio, box_noise_scale, num_queries, num_classes, noise=None):
    """-> (input_query_label [B,pad,num_classes], input_query_bbox [B,pad,7], attn_mask [T,T] bool, dn_meta)."""
    device = targets[0]["labels"].device
    batch_size = len(targets)
    known_num = [int(t["labels"].shape[0]) for t in targets]
    labels = torch.cat([t["labels"] for t in targets])
    boxes = torch.cat([t["gt_boxes"] for t in targets])
    batch_idx = torch.cat([torch.full((n,), i, dtype=torch.long, device=device) for i, n in enumerate(known_num)])
    total = int(labels.shape[0])
    reps = 2 * dn_number

    known_labels = labels.repeat(reps, 1).view(-1)
    known_bid = batch_idx.repeat(reps, 1).view(-1)
    known_bboxs = boxes.repeat(reps, 1)
    if noise is None:
        noise = draw_noise(reps * total, num_classes, device)
    else:
        noise = {k: v.to(device) for k, v in noise.items()}

    noised_labels = known_labels.clone()
    if label_noise_ratio > 0:
        chosen = noise["p_label"] < (label_noise_ratio * 0.5)
        noised_labels = torch.where(chosen, noise["new_label"].to(noised_labels.dtype), noised_labels)
    single_pad = int(max(known_num)) if known_num else 0
    pad_size = int(single_pad * reps)

    noised_boxes = known_bboxs.clone()
    if box_noise_scale > 0:
        corners = torch.cat((known_bboxs[:, :3] - known_bboxs[:, 3:6] / 2, known_bboxs[:, :3] + known_bboxs[:, 3:6] / 2,
                             known_bboxs[:, 6:]), dim=1)
        diff = torch.cat((known_bboxs[:, 3:6] / 2, known_bboxs[:, 3:6] / 2, torch.full_like(known_bboxs[:, 6:], 0.1)), dim=1)
        # negatives (odd repeats) get one extra half-size of displacement
        is_negative = (torch.arange(reps * total, device=device) // max(total, 1)) % 2 == 1
        rand_part = (noise["rand_part"] + is_negative[:, None].to(noise["rand_part"].dtype)) * noise["rand_sign"]
        corners = (corners + rand_part * diff * box_noise_scale).clamp(min=0.0, max=1.0)
        noised_boxes = torch.cat(((corners[:, :3] + corners[:, 3:6]) / 2, corners[:, 3:6] - corners[:, :3], corners[:, 6:]),
                                 dim=1)

    input_query_label = torch.zeros(batch_size, pad_size, num_classes, device=device)
    input_query_bbox = torch.zeros(batch_size, pad_size, 7, device=device)
    if total:
        local = torch.cat([torch.arange(n, device=device) for n in known_num])
        slot = torch.cat([local + single_pad * i for i in range(reps)]).long()
        input_query_label[known_bid, slot] = F.one_hot(noised_labels.long(), num_classes=num_classes).float()
        input_query_bbox[known_bid, slot] = noised_boxes

    tgt_size = pad_size + num_queries
    attn_mask = torch.zeros(tgt_size, tgt_size, dtype=torch.bool, device=device)
    attn_mask[pad_size:, :pad_size] = True   # matching queries cannot see the denoising queries
    attn_mask[:pad_size, pad_size:] = True   # denoising queries cannot see the matching queries
    grp = single_pad * 2
    for i in range(dn_number):               # denoising groups cannot see each other
        attn_mask[grp * i:grp * (i + 1), grp * (i + 1):pad_size] = True
        attn_mask[grp * i:grp * (i + 1), :grp * i] = True
    return input_query_label, input_query_bbox, attn_mask, {"pad_size": pad_size, "num_dn_group": dn_number}


def dn_post_process(outputs_class, outputs_coord, dn_meta, aux_loss):
    """Split the denoising part off the stacked decoder outputs ([layers, B, Q, .]) and park it in dn_meta."""
    if dn_meta and dn_meta["pad_size"] > 0:
        pad = dn_meta["pad_size"]
        known_cls, known_box = outputs_class[:, :, :pad], outputs_coord[:, :, :pad]
        out = {"pred_logits": known_cls[-1], "pred_boxes": known_box[-1]}
        if aux_loss:
            out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b} for a, b in zip(known_cls[:-1], known_box[:-1])]
        dn_meta["output_known_lbs_bboxes"] = out
        dn_meta["known_stack"] = (known_cls, known_box)   # the layer-stacked tensors, for the all-layers-at-once losses
        outputs_class, outputs_coord = outputs_class[:, :, pad:], outputs_coord[:, :, pad:]
    return outputs_class, outputs_coord
