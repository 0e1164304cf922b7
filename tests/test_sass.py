"""CPU tier: properties of the compiled sm_100a code that measurements depend on (checked with cuobjdump on the built library, no GPU).

* The library carries sm_100a code only.
* The warp-cooperative kernels (fused render kernel, fused trace kernel) contain no indirect branch: their dispatch on the warp's voted
  class used to compile to a jump table -- a constant-bank load + BRX behind the reduction, 17 times per tile, 2.3 % of the frame time
  (profiles/r02y_ab_nojumptable.log).
* Frames outside the GPU's own memory leave the CTA as bulk async copies: the render kernels contain UBLKCP (cp.async.bulk)."""
import re
import shutil
import subprocess

import pytest

from voxelraymarcher_b200 import api

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run(["cuobjdump", "-sass", api.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(line)
    return out, funcs


def test_library_is_sm_100a_only(sass):
    out, _ = sass
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs


def test_warp_cooperative_kernels_have_no_jump_table(sass):
    _, funcs = sass
    fused = [n for n in funcs if re.search(r"render_kernelILi0ELi0ELb[01]ELi2ELb[01]E", n) or "trace_fused_kernel" in n]
    assert len(fused) >= 4, fused   # statistics on/off x store form, + the trace kernels
    for n in fused:
        assert not any("BRX" in l or "JMX" in l for l in funcs[n]), n


def test_remote_frames_use_bulk_async_copies(sass):
    _, funcs = sass
    render = [n for n in funcs if "render_kernelI" in n and n.endswith("ELb0EEEvNS_10RenderArgsE")]   # the CTA-staged store form
    assert render
    for n in render:
        assert any("UBLKCP" in l for l in funcs[n]), n
