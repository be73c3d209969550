"""Host -> device staging for the entry points (the reference copies every minibatch synchronously in front of the
step, `.cuda(async=True)` on the compute stream: /root/reference/code/lib/model.py:220-225; its DataLoader uses
pin_memory=True, code/train.py:111-116).

`CudaPrefetcher(loader, device)` yields the loader's batches as device tensors; the copy of batch i+1 runs on a side
stream while the step of batch i computes (298 MB of int64 one-hot masks per 16-image batch is ~12 ms of PCIe time,
more than half a training step).  Every batch is still copied exactly once, from pinned host memory when the loader
provides it (non-pinned tensors are pinned here first).
"""
import numpy as np
import torch


def instance_label_map(instance_annotation, n_objects=None):
    """(h,w,n) uint8 per-instance masks, as the reference's LMDB stores them and its collate pads them
    (lib/dataset.py:36-58, :292-313) -> (h,w) uint8 label map, 255 = background.  1 byte per pixel instead of the
    8*MAX_N_OBJECTS bytes of the collate's int64 one-hot."""
    a = np.asarray(instance_annotation)
    if n_objects is not None:
        a = a[:, :, :int(n_objects)]
    if a.shape[2] == 0:
        return np.full(a.shape[:2], 255, dtype=np.uint8)
    lab = a.argmax(axis=2).astype(np.uint8)
    lab[a.max(axis=2) == 0] = 255
    return lab


def compact_collate(batch):
    """Collate for per-sample tuples (image (c,h,w) float tensor, semantic (h,w) uint8 class map, instance (h,w,n) uint8
    masks, n_objects) -- the items the reference's AlignCollate.__preprocess returns (lib/dataset.py:322) -- that emits
    what the kernels read: (images (b,c,h,w), sem (b,h,w) uint8 class map, ins (b,h,w) uint8 label map, n_objects).
    Replaces the int64 one-hot expansion of lib/dataset.py:354-376: 15 MB instead of 298 MB per 16-image batch."""
    images, sems, inss, nobj = zip(*batch)
    images = torch.stack([torch.as_tensor(i) for i in images])
    sem = torch.from_numpy(np.stack([np.asarray(s, dtype=np.uint8) for s in sems]))
    ins = torch.from_numpy(np.stack([instance_label_map(a, n) for a, n in zip(inss, nobj)]))
    return images, sem, ins, torch.IntTensor([int(n) for n in nobj])


def to_compact(sem_one_hot, ins_one_hot):
    """Reference-format HOST targets (one-hot (b,n_classes,h,w), (b,K,h,w)) -> (uint8 class map, uint8 label map)."""
    sem = sem_one_hot.max(1)[1].to(torch.uint8)
    val, idx = ins_one_hot.max(1)
    ins = idx.to(torch.uint8)
    ins[val == 0] = 255
    return sem, ins


class CudaPrefetcher(object):

    def __init__(self, loader, device, depth=1):
        self.loader = loader
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = max(1, int(depth))

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        out = []
        with torch.cuda.stream(self.stream):
            for t in batch:
                if torch.is_tensor(t):
                    if not t.is_cuda and not t.is_pinned():
                        t = t.pin_memory()
                    out.append(t.to(self.device, non_blocking=True))
                else:
                    out.append(t)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self):
        it = iter(self.loader)
        queue = []
        try:
            while len(queue) < self.depth:
                queue.append(self._stage(next(it)))
        except StopIteration:
            pass
        while queue:
            batch, ev = queue.pop(0)
            try:
                queue.append(self._stage(next(it)))     # next copy is in flight while the caller computes on `batch`
            except StopIteration:
                pass
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in batch:
                if torch.is_tensor(t):
                    t.record_stream(cur)                  # the side stream allocated it; the compute stream uses it
            yield batch
