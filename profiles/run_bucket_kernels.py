"""Launches each libdmlb bucket kernel (register path and TMA path) on a 1 GiB (cold, > L2) fp32 buffer through the C
ABI — the target of the `ncu --set full` capture (B200_PROFILING.md recipe).  Usage under gpurun:
    ncu --set full --clock-control none --import-source on -k regex:"stream_kernel|tma_kernel" -c 8 \
        -o gpurun_out/prof_bucket python profiles/run_bucket_kernels.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from dmlcloud_b200 import _native as N  # noqa: E402

lib = N.cuda_lib(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
src = torch.empty(n, dtype=torch.float32, device='cuda').normal_()
wire = torch.empty(n, dtype=torch.bfloat16, device='cuda')
st = N.stream_ptr()
N.check(lib.dmlb_bucket_pack_f32_bf16_regs(src.data_ptr(), wire.data_ptr(), n, 0.125, st))
N.check(lib.dmlb_bucket_pack_f32_bf16_tma(src.data_ptr(), wire.data_ptr(), n, 0.125, st))
N.check(lib.dmlb_bucket_unpack_bf16_f32_regs(wire.data_ptr(), src.data_ptr(), n, 1.0, None, st))
N.check(lib.dmlb_bucket_unpack_bf16_f32_tma(wire.data_ptr(), src.data_ptr(), n, 1.0, st))
N.check(lib.dmlb_bucket_scale_f32(src.data_ptr(), n, 1.0, st))
N.check(lib.dmlb_bucket_round_bf16_f32(src.data_ptr(), n, 1.0, None, st))
torch.cuda.synchronize()
print('done', N.launch_count())
