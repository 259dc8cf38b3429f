"""Generates tests/golden/proposals_*.npz from the REFERENCE's own gen_encoder_output_proposals
(/root/reference/models/richsem/utils.py:10-65), loaded by path in the build container.

    python tests/golden/make_golden_aux.py

Cases
  proposals_pad      2 images, 4 levels, C=8; image 0 unpadded, image 1 padded on the right / bottom
                     (a realistic DETR batch: masks are rectangles anchored top-left)
  proposals_learned  same, with learnedwh = (-1.3, 0.4) (two_stage_learn_wh)
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.aux_oracle import load_reference_proposals  # noqa: E402

HERE = Path(__file__).resolve().parent
SHAPES = [(20, 27), (10, 14), (5, 7), (3, 4)]


def rect_mask(shapes, fracs):
    """(N, S) bool: image i keeps the top-left (fh, fw) fraction of every level, the rest is padding."""
    rows = []
    for fh, fw in fracs:
        parts = []
        for h, w in shapes:
            m = torch.ones(h, w, dtype=torch.bool)
            m[: max(1, round(h * fh)), : max(1, round(w * fw))] = False
            parts.append(m.reshape(-1))
        rows.append(torch.cat(parts))
    return torch.stack(rows)


def main():
    ref = load_reference_proposals()
    assert ref is not None, "/root/reference is required to regenerate the golden vectors"
    g = torch.Generator().manual_seed(77)
    s = sum(h * w for h, w in SHAPES)
    memory = torch.randn(2, s, 8, generator=g)
    mask = rect_mask(SHAPES, [(1.0, 1.0), (0.7, 0.55)])
    shapes = torch.as_tensor(SHAPES, dtype=torch.long)
    for name, wh in (("proposals_pad", None), ("proposals_learned", torch.tensor([-1.3, 0.4]))):
        om, op = ref(memory, mask, shapes, wh)
        np.savez_compressed(HERE / f"{name}.npz", shapes=shapes.numpy(), memory=memory.numpy(), mask=mask.numpy(),
                            learnedwh=np.zeros(0, np.float32) if wh is None else wh.numpy(),
                            output_memory=om.numpy(), output_proposals=op.numpy())
        print(name, tuple(op.shape), "finite proposals:", int(torch.isfinite(op[..., 0]).sum()), "of", op.shape[0] * op.shape[1])


if __name__ == "__main__":
    main()
