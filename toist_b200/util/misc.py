"""Batch container used at the model boundary (mirrors the reference's util/misc.py:171-212 interface)."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor


class NestedTensor(object):
    """A padded image batch `tensors` [B, C, H, W] with its boolean padding mask `mask` [B, H, W] (True = padding)."""

    def __init__(self, tensors: Tensor, mask: Optional[Tensor]):
        self.tensors = tensors
        self.mask = mask

    def to(self, *args, **kwargs) -> "NestedTensor":
        t = self.tensors.to(*args, **kwargs)
        m = self.mask.to(*args, **kwargs) if self.mask is not None else None
        return type(self)(t, m)

    def decompose(self):
        return self.tensors, self.mask

    @classmethod
    def from_tensor_list(cls, tensor_list: List[Tensor], do_round: bool = False) -> "NestedTensor":
        """Pads every [C, h, w] image to the largest h / w of the list (optionally rounded up to a multiple of 128)."""
        if tensor_list[0].ndim != 3:
            raise ValueError("not supported")
        c = tensor_list[0].shape[0]
        h = max(int(t.shape[1]) for t in tensor_list)
        w = max(int(t.shape[2]) for t in tensor_list)
        if do_round:
            h = (h + 127) // 128 * 128
            w = (w + 127) // 128 * 128
        b = len(tensor_list)
        dtype, device = tensor_list[0].dtype, tensor_list[0].device
        if device.type == "cuda" and dtype == torch.float32:
            # one launch for the whole batch (the reference issues one copy and one mask fill per image)
            from .. import kernels as K

            tensor, mask_u8 = K.pad_batch_f32([t.contiguous() for t in tensor_list], h, w)
            return cls(tensor, mask_u8.view(torch.bool))
        tensor = torch.zeros((b, c, h, w), dtype=dtype, device=device)
        mask = torch.ones((b, h, w), dtype=torch.bool, device=device)
        for i, img in enumerate(tensor_list):
            tensor[i, :, : img.shape[1], : img.shape[2]].copy_(img)
            mask[i, : img.shape[1], : img.shape[2]] = False
        return cls(tensor, mask)

    @classmethod
    def from_uint8_list(cls, images: List[Tensor], mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225),
                        do_round: bool = False) -> "NestedTensor":
        """Decoded images as uint8 [h, w, 3] tensors (on the GPU, or on the host: copied through pinned staging) ->
        the normalised, padded fp32 batch + mask in ONE launch: ToTensor + Normalize (datasets/transforms.py:257-272,
        ImageNet statistics as in datasets/tdod.py:303) + from_tensor_list (util/misc.py:185-209).  Moves 1 byte per
        pixel and channel over PCIe instead of 4 and takes the per-image float work off the data-loader workers."""
        from .. import kernels as K

        dev = next((t.device for t in images if t.is_cuda), torch.device("cuda"))
        imgs = [h2d(t, dev).contiguous() for t in images]
        h = max(int(t.shape[0]) for t in imgs)
        w = max(int(t.shape[1]) for t in imgs)
        if do_round:
            h = (h + 127) // 128 * 128
            w = (w + 127) // 128 * 128
        tensor, mask_u8 = K.pad_normalize_u8(imgs, mean, std, h, w)
        return cls(tensor, mask_u8.view(torch.bool))

    def __repr__(self) -> str:
        return repr(self.tensors)


def h2d(t: torch.Tensor, device, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Host -> device copy that does not stall the host.  A copy from pageable memory makes the CUDA runtime drain the
    stream first (the host then waits for every kernel queued so far, once per small tensor, several times per step);
    staging through the caching pinned allocator keeps it asynchronous, so the host keeps issuing the next stage while
    the GPU works.  Tensors already on `device` are returned (converted) as they are."""
    device = torch.device(device)
    if t.device.type != "cpu" or device.type != "cuda":
        return t.to(device=device, dtype=dtype) if dtype is not None else t.to(device)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_pinned():
        t = t.contiguous().pin_memory()
    return t.to(device, non_blocking=True)



class Prefetcher:
    """Overlaps the host -> device copy of the NEXT batch with the current step (the reference copies
    synchronously at the top of every step, engine.py:51-58).  `batches` yields tuples / lists / dicts whose tensors
    live in pinned host memory; every batch is copied on a dedicated stream while the previous one is being consumed,
    and the consumer's stream is made to wait for exactly that copy.

        for samples, pmap in Prefetcher(loader, device):
            ...step...
    """

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.next = None
        self._fill()

    def _to(self, obj):
        if isinstance(obj, torch.Tensor):
            return obj.to(self.device, non_blocking=True) if obj.device.type == "cpu" else obj
        if isinstance(obj, NestedTensor):
            return NestedTensor(self._to(obj.tensors), self._to(obj.mask))
        if isinstance(obj, dict):
            return {k: self._to(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._to(v) for v in obj)
        return obj

    def _fill(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            self.next = None
            return
        with torch.cuda.stream(self.stream):
            dev = self._to(host)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.next = (dev, ev, host)

    def _record(self, obj, stream) -> None:
        if isinstance(obj, torch.Tensor):
            if obj.is_cuda:
                obj.record_stream(stream)
        elif isinstance(obj, NestedTensor):
            self._record(obj.tensors, stream)
            self._record(obj.mask, stream)
        elif isinstance(obj, dict):
            for v in obj.values():
                self._record(v, stream)
        elif isinstance(obj, (list, tuple)):
            for v in obj:
                self._record(v, stream)

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        dev, ev, _host = self.next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        self._record(dev, cur)  # allocated on the copy stream, used on the consumer's
        self._fill()            # the next batch's copy starts now and runs under this step
        return dev
