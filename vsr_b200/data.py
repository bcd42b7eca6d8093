"""Synthetic ACDC / DSB15-shaped cine data with pinned host staging (north_star item (d)).

The real datasets are NIfTI files read with nibabel (src/data/datasets/acdc_vsr_dataset.py:51-88);
here the same *contract* is produced from a seeded generator:
    {'lr_imgs': list of T tensors [1,h,w], 'hr_imgs': list of T tensors [1,r*h,r*w], 'index': i}
values are uint8-range intensities normalised with the dataset constants (transforms.py:154-168),
float32; LR = the reference's `Downscale` (k-space truncation + bicubic resize + round + clip,
acdc_preprocess.py:102-180) restated with numpy / cv2; training windows of `num_frames` frames with
temporal wrap-around (acdc_vsr_dataset.py:59-78); LR/HR-consistent random crop and flips
(transforms.py:321-450).
"""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .utils import DATASET_STATS

SHAPES = {"acdc": dict(hr=(128, 128), slices=10, frames=20), "dsb15": dict(hr=(256, 256), slices=12, frames=30)}


def synth_cine(h, w, frames, rng):
    """HR cine [frames,h,w]: moving anisotropic Gaussian blobs + smooth noise, periodic in time,
    rounded to integers in [0,255] (as the preprocessed data, acdc_preprocess.py:39-40)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((frames, h, w), np.float32)
    for _ in range(8):
        cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
        sy, sx = rng.uniform(0.04, 0.2) * h, rng.uniform(0.04, 0.2) * w
        amp, ph, mv = rng.uniform(40, 160), rng.uniform(0, 2 * np.pi), rng.uniform(0.0, 0.06)
        for t in range(frames):
            c = np.cos(2 * np.pi * t / frames + ph)
            s = 1.0 + 0.25 * c
            img[t] += amp * np.exp(-(((yy - cy - mv * h * c) / (sy * s)) ** 2 + ((xx - cx) / (sx * s)) ** 2) / 2)
    noise = rng.standard_normal((h // 4 + 1, w // 4 + 1)).astype(np.float32)
    noise = np.kron(noise, np.ones((4, 4), np.float32))[:h, :w] * 6.0
    return np.clip(np.round(img + noise[None]), 0, 255)


def downscale(img, r):
    """reference Downscale (acdc_preprocess.py:111-180) for one [h,w] image."""
    import cv2
    from numpy.fft import fftn, fftshift, ifftn, ifftshift
    img = np.asarray(img, dtype=np.float64)      # numpy 1.16 (the reference's pin) transforms in double precision
    k = fftshift(fftn(ifftshift(img), norm="ortho"))
    rect = np.zeros_like(k)
    kx, ky = k.shape[0] // 2, k.shape[1] // 2
    lx, ly = k.shape[0] // r, k.shape[1] // r
    rect[kx - lx // 2:kx + (lx - lx // 2), ky - ly // 2:ky + (ly - ly // 2)] = 1
    out = np.around(np.abs(fftshift(ifftn(ifftshift(rect * k), norm="ortho"))))
    out = cv2.resize(out, (img.shape[1] // r, img.shape[0] // r), interpolation=cv2.INTER_CUBIC)
    return np.clip(out.round(), 0, 255).astype(np.float32)


def lowpass_matrix(n, r):
    """The n x n complex matrix of the 1-D operator behind Downscale._truncate_kspace (acdc_preprocess.py:141-180):
    v -> fftshift(ifft(ifftshift(mask * fftshift(fft(ifftshift(v)))))) with mask = 1 on the centred n // r frequencies.
    The 2-D truncation of a frame X is  P_h @ X @ P_w.T  (all steps are separable).  [n, n, 2] float64 (re, im)."""
    from numpy.fft import fft, fftshift, ifft, ifftshift
    mask = np.zeros(n)
    k, l = n // 2, n // r
    mask[k - l // 2:k + (l - l // 2)] = 1
    eye = np.eye(n)
    P = fftshift(ifft(ifftshift(mask[:, None] * fftshift(fft(ifftshift(eye, axes=0), axis=0), axes=0), axes=0), axis=0), axes=0)
    return np.stack([P.real, P.imag], axis=-1).astype(np.float64)


_LOWPASS = {}


def downscale_device(hr, r, ops=None):
    """`downscale` for a stack of frames [n, h, w] already on the device (fp32, integer-valued): k-space truncation as two
    complex FP64 GEMMs per frame + integer-ratio bicubic + round + clip in three kernels (csrc/downscale.cu); bit-identical
    to the reference's Downscale (tests/golden/downscale.pt)."""
    if ops is None:
        from .ops import cuda_ops
        ops = cuda_ops()
    n, h, w = hr.shape
    mats = []
    for size in (h, w):
        key = (size, r, str(hr.device))
        if key not in _LOWPASS:
            _LOWPASS[key] = torch.from_numpy(lowpass_matrix(size, r)).to(hr.device)
        mats.append(_LOWPASS[key])
    lr = torch.empty(n, h // r, w // r, device=hr.device)
    ops.downscale(hr.contiguous().float(), r, mats[0], mats[1], lr)
    return lr


class SyntheticCineDataset(Dataset):
    """Args mirror AcdcVSRDataset (acdc_vsr_dataset.py:22): downscale_factor, num_frames,
    temporal_order, type ('train' | 'valid'); plus dataset ('acdc' | 'dsb15'), num_sequences,
    patch_size (LR crop for training, RandomCropPatch.size) and seed."""

    def __init__(self, downscale_factor, num_frames=5, temporal_order="last", type="train", dataset="acdc",
                 num_sequences=16, patch_size=(32, 32), seed=0, misr=False, device=None):
        if downscale_factor not in [2, 3, 4]:
            raise ValueError(f"The downscale factor should be 2, 3, 4. Got {downscale_factor}.")
        if temporal_order not in ["last", "middle"]:
            raise ValueError(f"The temporal order should be 'last' or 'middle'. Got {temporal_order}.")
        self.r, self.num_frames, self.temporal_order, self.type = downscale_factor, num_frames, temporal_order, type
        # misr=True: the AcdcMISRDataset contract (acdc_misr_dataset.py:46-88): every item, training or validation, is
        # a window of num_frames LR frames and ONE target, `hr_img` = the HR frame at num_frames // 2
        self.misr = misr
        self.patch = tuple(patch_size) if patch_size else None
        self.mean, self.std = DATASET_STATS[dataset]
        shape = SHAPES[dataset]
        rng = np.random.default_rng(seed)
        h, w = shape["hr"]
        h, w = h - h % self.r, w - w % self.r
        self.hr, self.lr = [], []
        for _ in range(num_sequences):
            cine = synth_cine(h, w, shape["frames"], rng)
            self.hr.append(cine)
            if device is None:
                self.lr.append(np.stack([downscale(f, self.r) for f in cine]))
        if device is not None:
            # `device`: the reference's Downscale for all frames of all sequences on the GPU (downscale_device: the same
            # integers as the host path, tests/test_elementwise_gpu.py) instead of one numpy FFT + cv2 call per frame
            hr_t = torch.from_numpy(np.stack(self.hr).astype(np.float32)).to(device)
            lr_t = downscale_device(hr_t.view(-1, h, w), self.r).view(num_sequences, shape["frames"], h // self.r, w // self.r)
            self.lr = list(lr_t.cpu().numpy())
        self.T = shape["frames"]
        self.rng = np.random.default_rng(seed + 1)
        self.data = [(s, t) for s in range(num_sequences) for t in range(self.T)] if type == "train" or misr \
            else [(s, None) for s in range(num_sequences)]

    def __len__(self):
        return len(self.data)

    def _window(self, t):
        n, T = self.num_frames, self.T
        if self.temporal_order == "last":
            start, end = t - n + 1, t + 1
        else:
            start, end = t - (n - 1) // 2, t + ((n - 1) - (n - 1) // 2) + 1
        return [i % T for i in range(start, end)]

    def draw(self, index):
        """The random decisions of one item, in the order __getitem__ consumes the generator:
        (sequence, frame indices, flip_x, flip_y, y0, x0, patch_h, patch_w) — the crop is taken from the flipped image
        (transforms.py:321-450: RandomHorizontalFlip, RandomVerticalFlip, then RandomCropPatch)."""
        s, t = self.data[index]
        idx = self._window(t) if self.type == "train" or self.misr else list(range(self.T))
        h, w = self.lr[s].shape[1:]
        fx = fy = False
        y0 = x0 = 0
        ph, pw = h, w
        if self.type == "train":
            fx = bool(self.rng.random() < 0.5)
            fy = bool(self.rng.random() < 0.5)
            if self.patch:
                ph, pw = self.patch
                y0 = int(self.rng.integers(0, h - ph + 1))
                x0 = int(self.rng.integers(0, w - pw + 1))
        return s, idx, fx, fy, y0, x0, ph, pw

    def __getitem__(self, index):
        s, idx, fx, fy, y0, x0, ph, pw = self.draw(index)
        lr, hr = self.lr[s][idx], self.hr[s][idx]
        if fx:
            lr, hr = lr[:, :, ::-1], hr[:, :, ::-1]
        if fy:
            lr, hr = lr[:, ::-1], hr[:, ::-1]
        lr = lr[:, y0:y0 + ph, x0:x0 + pw]
        hr = hr[:, y0 * self.r:(y0 + ph) * self.r, x0 * self.r:(x0 + pw) * self.r]
        norm = lambda a: torch.from_numpy(((np.ascontiguousarray(a) - self.mean) / self.std).astype(np.float32))
        lr, hr = norm(lr), norm(hr)
        if self.misr:
            n = self.num_frames
            return {"lr_imgs": [f.unsqueeze(0) for f in lr], "hr_img": hr[n // 2 if n % 2 == 1 else n // 2 - 1].unsqueeze(0),
                    "index": index}
        return {"lr_imgs": [f.unsqueeze(0) for f in lr], "hr_imgs": [f.unsqueeze(0) for f in hr], "index": index}


class Dataloader(DataLoader):
    """Same constructor keywords as the reference Dataloader (src/data/dataloader.py:6-53), with
    pinned host memory on by default so H2D copies can be asynchronous."""

    def __init__(self, dataset, batch_size=1, shuffle=False, sampler=None, batch_sampler=None, num_workers=0,
                 collate_fn=None, pin_memory=True, drop_last=False, timeout=0, worker_init_fn=None):
        if worker_init_fn is None:
            worker_init_fn = self._default_worker_init_fn
        kw = dict(dataset=dataset, batch_size=batch_size, shuffle=shuffle, sampler=sampler,
                  batch_sampler=batch_sampler, num_workers=num_workers,
                  pin_memory=pin_memory and torch.cuda.is_available(), drop_last=drop_last, timeout=timeout,
                  worker_init_fn=worker_init_fn)
        if collate_fn is not None:
            kw["collate_fn"] = collate_fn
        super().__init__(**kw)

    @staticmethod
    def _default_worker_init_fn(worker_id):
        np.random.seed((np.random.get_state()[1][0] + worker_id) % (2 ** 32))


class ShardedSampler(torch.utils.data.Sampler):
    """Rank-aware, seed-deterministic sampler for the data-parallel trainer.  The reference reseeds
    numpy once per epoch (base_trainer.py:54) and shuffles inside its single process; here every rank draws the SAME
    permutation from the epoch's seed (`set_epoch_seed`, called by VSRTrainer with that very seed) and takes every
    world-th index starting at its rank.  The index list is padded by wrap-around to a multiple of the world size
    (as torch's DistributedSampler does), so all ranks see the same number of batches - a rank that ran out early
    would leave the others blocked in the gradient all-reduce."""

    def __init__(self, n, rank, world, shuffle):
        self.n, self.rank, self.world, self.shuffle = int(n), int(rank), int(world), bool(shuffle)
        self.seed = 0

    def set_epoch_seed(self, seed):
        self.seed = int(seed) % (2 ** 32)

    def indices(self):
        order = np.random.RandomState(self.seed).permutation(self.n) if self.shuffle else np.arange(self.n)
        total = -(-self.n // self.world) * self.world
        if total > self.n:
            order = np.concatenate([order, order[:total - self.n]])
        return order[self.rank::self.world]

    def __iter__(self):
        return iter(self.indices().tolist())

    def __len__(self):
        return -(-self.n // self.world)


def shard_loader(loader, rank, world):
    """the same loader restricted to this rank's shard; returns (loader, sampler or None).  torch DataLoaders are
    rebuilt around a ShardedSampler (same dataset, batch size, workers, collate, pinning); DeviceCineLoader takes the
    sampler directly; anything else is returned unchanged (the trainer then checks the batch counts across ranks)."""
    if world <= 1:
        return loader, None
    if isinstance(loader, DeviceCineLoader):
        sampler = ShardedSampler(len(loader.dataset), rank, world, loader.shuffle)
        loader.sampler = sampler
        return loader, sampler
    if isinstance(loader, DataLoader):
        if isinstance(loader.sampler, (ShardedSampler, torch.utils.data.distributed.DistributedSampler)):
            return loader, loader.sampler if isinstance(loader.sampler, ShardedSampler) else None
        if loader.batch_sampler is None or loader.batch_size is None:
            return loader, None
        shuffle = isinstance(loader.sampler, torch.utils.data.RandomSampler)
        sampler = ShardedSampler(len(loader.dataset), rank, world, shuffle)
        new = loader.__class__.__new__(loader.__class__)
        DataLoader.__init__(new, loader.dataset, batch_size=loader.batch_size, sampler=sampler,
                            num_workers=loader.num_workers, collate_fn=loader.collate_fn, pin_memory=loader.pin_memory,
                            drop_last=loader.drop_last, timeout=loader.timeout, worker_init_fn=loader.worker_init_fn)
        return new, sampler
    return loader, None


class DeviceStager:
    """Wraps a batch iterator: copies each batch from pinned host memory to the device on a side
    stream one step ahead of the consumer (replaces the blocking, pageable `tensor.to(device)` of
    base_trainer.py:146-161)."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def _move(self, obj):
        if isinstance(obj, torch.Tensor):
            if obj.device.type == self.device.type and self.device.index in (None, obj.device.index):
                # DeviceCineLoader batches are already there
                return obj
            src = obj if obj.is_pinned() else obj.pin_memory()
            return src.to(self.device, non_blocking=True)
        if isinstance(obj, dict):
            return {k: self._move(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            frames = list(obj)
            if (self.device.type == "cuda" and len(frames) > 1 and all(isinstance(v, torch.Tensor) and not v.is_cuda for v in frames)
                    and all(v.shape == frames[0].shape and v.dtype == frames[0].dtype for v in frames)):
                # the T frames of a sequence land in ONE [T, ...] device buffer (a copy per frame, no host-side stack):
                # the fused step then takes all frames with one device copy and one loss / metric launch
                buf = torch.empty(len(frames), *frames[0].shape, dtype=frames[0].dtype, device=self.device)
                for dst, v in zip(buf.unbind(0), frames):
                    dst.copy_(v if v.is_pinned() else v.pin_memory(), non_blocking=True)
                return type(obj)(buf.unbind(0))
            return type(obj)(self._move(v) for v in obj)
        return obj

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        it = iter(self.loader)

        def fetch():
            try:
                host = next(it)
            except StopIteration:
                return None
            with torch.cuda.stream(self.stream):
                dev = self._move(host)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            return dev, ev

        def claim(obj, stream):
            # the staged tensors were allocated on the side stream: tell the caching allocator that the consumer's
            # stream uses them, so that their blocks are not handed to a later prefetch (whose H2D copy has no
            # dependency on the consumer) before the queued kernels have read them
            if isinstance(obj, torch.Tensor):
                if obj.is_cuda:
                    obj.record_stream(stream)
            elif isinstance(obj, dict):
                for v in obj.values():
                    claim(v, stream)
            elif isinstance(obj, (list, tuple)):
                for v in obj:
                    claim(v, stream)

        nxt = fetch()
        while nxt is not None:
            dev, ev = nxt
            nxt = fetch()
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            claim(dev, cur)
            yield dev


class DeviceCineLoader:
    """The same batches as `Dataloader(SyntheticCineDataset(...))` without the host in the loop (SURVEY §8f rank 3,
    the augmentation half): the cine volumes live in device memory, and one kernel per batch (`vsr_cine_gather`) does
    what the reference does per item on the host — temporal window (acdc_vsr_dataset.py:59-78), RandomHorizontalFlip /
    RandomVerticalFlip / RandomCropPatch (transforms.py:321-450), Normalize (:154-168), ToTensor and the default
    collate — writing the collated `[N,1,h,w]` frames directly.  The random decisions come from the dataset's own
    generator in the dataset's own order (`SyntheticCineDataset.draw`), so a batch is bit-identical to the host
    loader's batch of the same items.  Sequences of one dataset share a shape."""

    def __init__(self, dataset, device, batch_size=1, shuffle=False, drop_last=False, ops=None):
        self.dataset, self.device, self.batch_size = dataset, torch.device(device), batch_size
        self.shuffle, self.drop_last, self._ops = shuffle, drop_last, ops
        self.sampler = None          # a ShardedSampler under the data-parallel trainer (shard_loader)
        self.lr = torch.from_numpy(np.stack(dataset.lr).astype(np.float32)).to(self.device)      # [S, T, h, w]
        self.hr = torch.from_numpy(np.stack(dataset.hr).astype(np.float32)).to(self.device)      # [S, T, rh, rw]

    def __len__(self):
        n = len(self.sampler) if self.sampler is not None else len(self.dataset)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def _backend(self):
        if self._ops is not None:
            return self._ops
        from .ops import cuda_ops
        return cuda_ops()

    def batch(self, indices):
        ds, ops = self.dataset, self._backend()
        draws = [ds.draw(int(i)) for i in indices]
        nf = len(draws[0][1])
        ph, pw = draws[0][6], draws[0][7]
        n = len(draws)
        # per item: sequence, flip_x, flip_y, y0, x0, then the nf frame indices
        tab = torch.tensor([[d[0], int(d[2]), int(d[3]), d[4], d[5]] + list(d[1]) for d in draws], dtype=torch.int32)
        tab = tab.to(self.device, non_blocking=True)
        hr_frames = [nf // 2 if nf % 2 == 1 else nf // 2 - 1] if ds.misr else list(range(nf))
        lr = torch.empty(nf, n, 1, ph, pw, device=self.device)
        hr = torch.empty(len(hr_frames), n, 1, ph * ds.r, pw * ds.r, device=self.device)
        ops.cine_gather(self.lr, tab, 1, 0, nf, ds.mean, ds.std, lr)
        ops.cine_gather(self.hr, tab, ds.r, hr_frames[0], len(hr_frames), ds.mean, ds.std, hr)
        out = {"lr_imgs": list(lr.unbind(0)), "index": torch.as_tensor(list(indices))}
        if ds.misr:
            out["hr_img"] = hr[0]
        else:
            out["hr_imgs"] = list(hr.unbind(0))
        return out

    def __iter__(self):
        if self.sampler is not None:
            order = self.sampler.indices()
        else:
            n = len(self.dataset)
            order = np.random.permutation(n) if self.shuffle else np.arange(n)
        for b in range(len(self)):
            yield self.batch(order[b * self.batch_size:(b + 1) * self.batch_size])
