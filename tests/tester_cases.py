"""Inputs for the Tester.test_epoch composition tests (shared by the CPU harness test and the GPU drop-in test)."""
import torch

C, H, W, T, STEPS = 20, 32, 256, 4, 3


def make_case(branch: str, seed: int = 0):
    """(batches, model_outputs): loader batches of 5 CPU tensors and the head outputs in model-call order.
    The true class gets a +2 logit bonus so predictions correlate with the labels (non-trivial IoU / AUROC)."""
    g = torch.Generator().manual_seed(seed)
    batches = []
    for _ in range(STEPS):
        lab = torch.randint(0, C, (1, 1, H, W), generator=g)
        batches.append((torch.rand(1, 1, H, W, generator=g), torch.rand(1, 1, H, W, generator=g),
                        torch.randn(1, 3, H, W, generator=g), torch.randn(1, 3, H, W, generator=g), lab))
    per_step = T if branch == "mc" else 1
    ch = C if branch == "mc" else C + 1
    outs = []
    for i in range(STEPS * per_step):
        o = torch.randn(1, ch, H, W, generator=g) * 3.0
        lab = batches[i // per_step][4][:, 0]
        o[:, :C].scatter_add_(1, lab[:, None], torch.full((1, 1, H, W), 2.0))
        outs.append(o)
    return batches, outs
