"""mslesseg_b200 - B200-native voxel path of YOLO-MSLesSeg (enhance->slice, recon->consensus->eval).

`ops`      batched device-tensor API over the C ABI (include/mslesseg.h, libmslesseg.so)
`compat`   drop-in mirrors of the reference's Python call sites (same names / arguments / errors)
`metrics`  host-side float64 metric formulas fed by the device counts
`dist`     patient sharding and the NCCL all-reduce of the count table
`synthetic` seeded MSLesSeg-shaped cohorts for tests and bench.py
"""
__version__ = "0.1.0"
