"""denormalize (reference: src/utils.py:1-20) and dataset constants."""
import torch

DATASET_STATS = {"acdc": (54.089, 48.084), "dsb15": (51.193, 52.671)}   # utils.py:13-16


def denormalize(imgs, dataset):
    """(imgs * std + mean).round().clamp(0, 255) with the dataset's constants (utils.py:18-20).
    Kept for callers that need the denormalised images themselves (e.g. loggers); the metric
    kernels fuse this step (FusedPSNR / FusedSSIM, or PSNR/SSIM(dataset=...))."""
    if dataset not in DATASET_STATS:
        raise ValueError(f"The name of the dataset should be 'acdc' or 'dsb15'. Got {dataset}.")
    mean, std = DATASET_STATS[dataset]
    return (imgs.clone() * std + mean).round().clamp(0, 255)
