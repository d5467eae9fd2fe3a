"""Mirror of the reference's ``transformer/utils.py`` (:1-9)."""
import torch

DEVICE = 'cuda' if torch.cuda.is_available() else 'cpu'


def init_device():
    """transformer/utils.py:3-5 -- sets the module-global DEVICE."""
    global DEVICE
    DEVICE = 'cuda' if torch.cuda.is_available() else 'cpu'


def count_parameters(model):
    """transformer/utils.py:8-9."""
    return sum([p.numel() for p in model.parameters() if p.requires_grad])
