"""Input side of the hot path: reader for the reference's BATCHED on-disk format with a pinned-memory
prefetch ring.

The reference stores one `torch.save`d float32 tensor of shape (B, 7, H, W) per training batch
(dmmfods/utils/Dense_U_Net_lidar_helper.py:653-728, layout :717-719) and `WaymoDataset.get_batch`
(dmmfods/datasets/WaymoData.py:87-103) splits it into image = [:, :3], lidar = [:, 3:4], heat maps = [:, 4:].
`BatchFileRing` yields exactly those three tensors, already in PINNED host memory (so that `Trainer.prefetch` can copy them
asynchronously over PCIe while the previous step computes); a background thread reads and splits `depth` batches ahead.
The file list is the reference's: every file of `<root>/<mode>/<subdir>/` except the `labels` entry (WaymoData.py:45-52).
"""
import os
import queue
import threading

import torch


def list_batch_files(root, mode):
    """relative paths `<mode>/<subdir>/<batch>` of a batched dataset, in the reference's crawl order (WaymoData.py:45-52)."""
    files = []
    for subdir in os.listdir(os.path.join(root, mode)):
        entries = os.listdir(os.path.join(root, mode, subdir))
        if "labels" in entries:
            entries.remove("labels")
        files += [os.path.join(mode, subdir, b) for b in entries]
    return files


def split_batch(batch):
    """(B,7,H,W) -> image (B,3,H,W), lidar (B,1,H,W), heat maps (B,3,H,W) views (WaymoData.py:98-101)."""
    if batch.dim() != 4 or batch.shape[1] != 7:
        raise ValueError("expected a (B, 7, H, W) batch tensor, got %s" % (tuple(batch.shape),))
    return batch[:, :3, :, :], batch[:, 3, :, :].unsqueeze(1), batch[:, 4:, :, :]


def mark_async_read(tensors, event):
    """tag host tensors with the CUDA event recorded after an asynchronous (non_blocking) device copy that reads them."""
    for t in tensors:
        if t is not None and not t.is_cuda:
            t._dmm_read_event = event


def wait_async_reads(t):
    """block until the asynchronous device copies that were reading host tensor `t` (mark_async_read) have completed."""
    ev = getattr(t, "_dmm_read_event", None)
    if ev is not None:
        ev.synchronize()
        t._dmm_read_event = None


class BatchFileRing:
    """iterate over batch files as (image, lidar, heat_maps) pinned float32 tensors, `depth` batches read ahead.

    Every slot of the ring owns its pinned buffers; a yielded triple stays valid until TWO further batches have been taken
    from the iterator (one step in flight + one prefetch).  The reader thread fills a slot BEFORE it blocks on the queue of
    `depth - 1` finished batches, so depth + 2 slots are needed: the slot it overwrites for batch n held batch n - depth - 2,
    and at that moment at least n - (depth - 1) batches have been taken.

    Slot reuse is additionally tied to COPY COMPLETION: a consumer that reads a yielded tensor asynchronously (the
    non-blocking host -> device copies of `Trainer.prefetch` / `Trainer.step`) tags it with the CUDA event recorded after
    the copy (`mark_async_read`), and the reader thread waits for that event before it overwrites the slot - a host that
    runs ahead of the GPU (CUDA-graph steps, page-cached files) can therefore never corrupt an in-flight batch."""

    def __init__(self, root, files, depth=3, pin=None, epochs=1):
        self.root, self.files, self.depth, self.epochs = root, list(files), max(2, int(depth)), epochs
        self.pin = torch.cuda.is_available() if pin is None else bool(pin)
        self._slots = [None] * (self.depth + 2)

    def __len__(self):
        return len(self.files) * self.epochs

    def _load(self, rel, slot):
        batch = torch.load(os.path.join(self.root, rel))
        parts = split_batch(batch.float() if batch.dtype != torch.float32 else batch)
        bufs = self._slots[slot]
        if bufs is None or any(b.shape != p.shape for b, p in zip(bufs, parts)):
            bufs = tuple(torch.empty(p.shape, dtype=torch.float32, pin_memory=self.pin) for p in parts)
            self._slots[slot] = bufs
        for b, p in zip(bufs, parts):
            wait_async_reads(b)
            b.copy_(p)
        return bufs

    def __iter__(self):
        q = queue.Queue(maxsize=self.depth - 1)
        stop = threading.Event()

        def reader():
            try:
                n = 0
                for _ in range(self.epochs):
                    for rel in self.files:
                        if stop.is_set():
                            return
                        q.put(self._load(rel, n % len(self._slots)))
                        n += 1
                q.put(None)
            except Exception as e:      # noqa: BLE001  (re-raised in the consumer)
                q.put(e)
        t = threading.Thread(target=reader, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, Exception):
                    raise item
                yield item
        finally:
            stop.set()
            while not q.empty():
                q.get_nowait()
