"""PyTorch/cuDNN rebuild of the reference's `dpc/nets`: library kernels only --
the networks around the renderer are dense conv/GEMM work for cuDNN/cuBLAS."""
