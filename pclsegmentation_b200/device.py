"""PyTorch as plumbing: device buffers, streams and pinned staging for the C-ABI calls."""
import numpy as np
import torch

from ._lib import PclsError


def require_cuda():
  if not torch.cuda.is_available():
    raise PclsError("no CUDA device visible: pclsegmentation_b200 has no CPU fallback")
  return torch.device("cuda", torch.cuda.current_device())


def ptr(t):
  """Device pointer of a torch tensor (None -> NULL)."""
  if t is None:
    return None
  if not t.is_cuda:
    raise PclsError("expected a CUDA tensor")
  if not t.is_contiguous():
    raise PclsError("expected a contiguous tensor")
  return t.data_ptr()


def stream_handle():
  return torch.cuda.current_stream().cuda_stream


def to_device(x, dtype, pinned_cache=None, key=None):
  """numpy array / torch tensor -> contiguous CUDA tensor of `dtype` (async H2D through pinned memory)."""
  dev = require_cuda()
  if torch.is_tensor(x):
    if x.is_cuda:
      return x.to(dtype).contiguous()
    return x.to(dtype).contiguous().pin_memory().to(dev, non_blocking=True)
  a = np.ascontiguousarray(x)
  t = torch.from_numpy(a)
  if t.dtype != dtype:
    t = t.to(dtype)
  if pinned_cache is not None:
    buf = pinned_cache.get(key)
    if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
      buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
      pinned_cache[key] = buf
    buf.copy_(t)
    t = buf
  else:
    t = t.pin_memory()
  return t.to(dev, non_blocking=True)


class DeviceTensor(torch.Tensor):
  """A CUDA tensor whose ``.numpy()`` performs the device->host copy, so reference code written against TF
  eager tensors (``predictions.numpy()[0]``, inference.py:78) works unchanged."""

  def numpy(self):  # noqa: D401
    return self.detach().as_subclass(torch.Tensor).cpu().numpy()


def wrap(t):
  return t.as_subclass(DeviceTensor)
