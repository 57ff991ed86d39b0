"""Host-side mirror of the reference's `dpc/util` modules that sit on the projection
hot path: point_cloud, drc, gauss_kernel, quaternion, camera, config."""
