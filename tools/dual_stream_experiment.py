"""Experiment: one batch-32 forward per step vs two batch-16 forwards on two streams (two nets, two arenas), so that the
tail / prologue of one stream's layer is filled by the other stream's layer.  Prints frames/s of both forms."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


class A:
  conv_impl = use_graph = micro_batch = None
  opt = []


def main():
  dev = torch.device("cuda", 0)
  torch.cuda.set_device(dev)
  B = 32
  name, mc, m_full, _ = bench.build_model(bench.DEFAULT_WORKLOAD, A, B)
  H, W, NC = mc.ZENITH_LEVEL, mc.AZIMUTH_LEVEL, mc.NUM_CLASS
  raw = [torch.from_numpy(bench.synth_raw(1234 + i, B, H, W)).to(dev) for i in range(4)]
  outs = [{"predictions": torch.empty((B, H, W), dtype=torch.int32, device=dev),
           "probabilities": torch.empty((B, H, W, NC), dtype=torch.float32, device=dev)} for _ in range(2)]

  def time_it(step, n=30, warm=8):
    for i in range(warm):
      step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
      step(i)
    b.record()
    torch.cuda.synchronize()
    return B * n / (a.elapsed_time(b) / 1e3)

  def full(i):
    m_full.forward_device(raw[i % 4], None, mean=mc.INPUT_MEAN, std=mc.INPUT_STD, out=outs[i % 2])
  print("one stream, batch 32: %.0f frames/s" % time_it(full))

  for parts in (2, 4):
    hb = B // parts
    models = [bench.build_model(bench.DEFAULT_WORKLOAD, A, hb)[2] for _ in range(parts)]
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    views = [[{k: v[p * hb:(p + 1) * hb] for k, v in o.items()} for p in range(parts)] for o in outs]
    rviews = [[r[p * hb:(p + 1) * hb] for p in range(parts)] for r in raw]
    main_s = torch.cuda.current_stream()

    def split(i):
      ev = torch.cuda.Event()
      ev.record(main_s)
      for p in range(parts):
        streams[p].wait_event(ev)
        with torch.cuda.stream(streams[p]):
          models[p].forward_device(rviews[i % 4][p], None, mean=mc.INPUT_MEAN, std=mc.INPUT_STD, out=views[i % 2][p])
          e2 = torch.cuda.Event()
          e2.record(streams[p])
        main_s.wait_event(e2)
    print("%d streams, batch %d each: %.0f frames/s" % (parts, hb, time_it(split)))


if __name__ == "__main__":
  main()
