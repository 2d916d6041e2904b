"""lama-b200: the particle-filter SLAM hot path of iris-ua/iris_lama (LaMa) as sm_100a CUDA kernels behind a C-ABI.

    api          ctypes mirror of include/lama_b200.h with the reference's class / method names
                 (PFSlam2D, Slam2D, LidarOdometry2D, Loc2D, DynamicDistanceMap)
    distributed  one process per GPU over torch.distributed (ShardedPFSlam2D)
    sdm          numpy mirror of the reference's on-disk map format (.sdm)
    synth        seeded synthetic worlds / scans / odometry for tests and bench.py

Nothing is imported eagerly: `api` loads iris_lama_b200/liblama_b200.so (built by `__graft_entry__.build()`), which
needs a CUDA device at run time -- there is no CPU fallback.
"""
