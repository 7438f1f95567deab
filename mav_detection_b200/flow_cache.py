"""The `.flo` side of the flow-provider seam (next-row f2): Farneback flow computed on the device is written in the
on-disk format every other consumer of the reference expects (/root/reference/src/utils.py:204-257), under the path
convention Dataset.get_flow_uv reads (/root/reference/src/datasets/dataset.py:205-212), and served back through the
same accessor.

    cache = FloCache(f'{img_path}/output/inference/run.epoch-0-flow-field')
    cache.put_batch(first_index, flow)          # (n, H, W, 2) float32 CUDA tensor or host array
    flow_i = cache.get_flow_uv(i)               # (H, W, 2) float32, == utils.read_flow of the file

A `.flo` file is a 12-byte header (float32 tag 202021.25, int32 width, int32 height) followed by the dense (H, W, 2)
float32 field — exactly the layout mavd_farneback writes — so a batch goes device -> pinned host staging (one
asynchronous copy per batch, two staging buffers so that the copy of batch k+1 overlaps the file writes of batch k)
-> header + payload per frame, with no re-arrangement on the host."""
from __future__ import annotations

import os
from typing import Any, Optional

import numpy as np

from . import utils

TAG = np.array([utils.TAG_FLOAT], np.float32).tobytes()


def flo_path(directory: str, i: int) -> str:
    """dataset.py:211: f'{img_path}/output/inference/run.epoch-0-flow-field/{i:06d}.flo'."""
    return os.path.join(directory, '%06d.flo' % i)


class FloCache:
    def __init__(self, directory: str) -> None:
        self.directory = directory
        utils.create_if_not_exists(directory)
        self._staging = [None, None]
        self._pending = None          # (first_index, n, staging array, cuda event)
        self._turn = 0

    # -- writing ---------------------------------------------------------------------------------
    def _write_frames(self, first_index: int, flows: np.ndarray) -> None:
        n, h, w = flows.shape[:3]
        header = TAG + np.array([w, h], np.int32).tobytes()
        for k in range(n):
            with open(flo_path(self.directory, first_index + k), 'wb') as f:
                f.write(header)
                flows[k].tofile(f)

    def flush(self) -> None:
        """Writes the batch whose device->host copy is still in flight."""
        if self._pending is not None:
            first, n, stage, event = self._pending
            self._pending = None
            event.synchronize()
            self._write_frames(first, stage[:n])

    def put_batch(self, first_index: int, flow: Any) -> None:
        """Frames first_index .. first_index + n - 1 of a (n, H, W, 2) float32 batch."""
        if isinstance(flow, np.ndarray):
            if flow.dtype != np.float32 or flow.ndim != 4 or flow.shape[3] != 2:
                raise ValueError('flow must be float32 (n, H, W, 2)')
            self.flush()
            self._write_frames(first_index, np.ascontiguousarray(flow))
            return
        import torch
        if flow.dtype != torch.float32 or flow.dim() != 4 or flow.shape[3] != 2 or not flow.is_cuda:
            raise ValueError('flow must be a float32 (n, H, W, 2) CUDA tensor or host array')
        flow = flow.contiguous()
        t = self._turn
        self._turn ^= 1
        stage = self._staging[t]
        if stage is None or stage.shape[0] < flow.shape[0] or tuple(stage.shape[1:]) != tuple(flow.shape[1:]):
            stage = torch.empty(tuple(flow.shape), dtype=torch.float32, pin_memory=True)
            self._staging[t] = stage
        stage[:flow.shape[0]].copy_(flow, non_blocking=True)
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(flow.device))
        previous = self._pending
        self._pending = None
        if previous is not None:                  # write batch k while batch k+1 is on its way
            first, n, st, ev = previous
            ev.synchronize()
            self._write_frames(first, st[:n])
        self._pending = (first_index, int(flow.shape[0]), stage.numpy(), event)

    # -- reading (the Dataset.get_flow_uv seam) -----------------------------------------------------
    def has(self, i: int) -> bool:
        self.flush()
        return os.path.exists(flo_path(self.directory, i))

    def get_flow_uv(self, i: int) -> Optional[np.ndarray]:
        """dataset.py:205-212: the content of frame i's .flo file, (H, W, 2) float32."""
        self.flush()
        return utils.read_flow(flo_path(self.directory, i))


class CachedFlowDataset:
    """Wraps any object with the reference's Dataset interface and serves get_flow_uv from a FloCache (every other
    attribute is forwarded), so that a sequence whose flow was computed once on the device can be re-run through
    Processor(flow_source='dataset') exactly like a dataset with FlowNet2 .flo files."""

    def __init__(self, dataset: Any, cache: FloCache) -> None:
        self._dataset = dataset
        self.flow_cache = cache

    def get_flow_uv(self, i: int) -> np.ndarray:
        return self.flow_cache.get_flow_uv(i)

    def __getattr__(self, name: str) -> Any:
        return getattr(self._dataset, name)
