"""b200vt — B200-native (sm_100a) attention hot path for VideoTuna-style video denoisers.

Public surface (mirrors the reference callables):
    b200vt.ops        torch.library ops over the C ABI (libb200vt.so)
    b200vt.functional reference-signature functions: hunyuan `attention`, wan `flash_attention`,
                      lvdm `CrossAttention.forward`, modulate/gate/norm helpers
    b200vt.blocks     block-level drop-in forwards (Hunyuan / Wan / lvdm / diffusers blocks) over the fused row kernels
    b200vt.sp         Ulysses sequence parallelism (NCCL all-to-all, fused output exchange, host-buffer pipelines)
    b200vt.patch      patch_videotuna() / patch_blocks() / patch_sp(): rebind the reference's hooks to the functions above
    b200vt.xfuser_shim  stand-in for the xfuser names the reference's sequence-parallel entry points import
    b200vt.graph      whole-step CUDA graph capture for the launch-bound lvdm UNet

There is no CPU fallback and no alternative backend: calling an op without the CUDA library raises.
"""
__version__ = "0.1.0"
