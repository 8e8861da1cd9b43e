"""oracle/ — CPU restatements of the reference algorithms and bindings to the reference's own code.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
reference arm.  Nothing under gs_icp_slam_b200/, diff_gaussian_rasterization/, pygicp/ or simple_knn/
imports this package (tests/test_abi.py::test_product_does_not_import_oracle enforces it).
"""
