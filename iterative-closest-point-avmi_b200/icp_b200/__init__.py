"""icp_b200 -- Python host layer of the B200-native ICP / occupancy hot path.

``icp_b200.api`` wraps the C ABI of libicp_b200.so (include/icp_b200.h);
``icp_b200.synth`` generates the deterministic synthetic lidar inputs;
``icp_b200.dist`` shards batches over one-process-per-GPU ranks.
The drop-in replacements for the reference's ``utilities.icp`` /
``utilities.mapping`` live next to this package in ``utilities/``.
"""
