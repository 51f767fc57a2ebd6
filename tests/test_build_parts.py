"""conv_tc.cu is compiled once per PART (-DPCLS_TC_PART=k): the build script's part count must match the parts the source
defines, otherwise a dispatcher would reference an instantiation no translation unit emits (link error at best)."""
import os
import re

from pclsegmentation_b200 import build


def test_tc_parts_match_source():
  src = open(os.path.join(build.CSRC, "conv_tc.cu")).read()
  parts = sorted({int(m) for m in re.findall(r"#if PCLS_TC_PART == (\d+)", src)})
  assert parts == list(range(1, build.TC_PARTS)), (parts, build.TC_PARTS)      # part 0 = host code + special kernels
  assert "#if PCLS_TC_PART <= 0" in src
  tc = [s for s in build.SOURCES if isinstance(s, tuple)]
  assert tc == [("conv_tc.cu", k) for k in range(build.TC_PARTS)]


def test_every_source_file_is_built():
  built = {s[0] if isinstance(s, tuple) else s for s in build.SOURCES}
  on_disk = {f for f in os.listdir(build.CSRC) if f.endswith(".cu")}
  assert built == on_disk, (sorted(built ^ on_disk))
