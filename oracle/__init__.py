"""Oracle package: TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Nothing under
bcftools_b200/ imports this package (tests/test_no_oracle_in_product.py enforces it).
"""
