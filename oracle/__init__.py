"""Test infrastructure: CPU oracle, seeded inputs and golden-vector tooling for the MHAda hot path.

Nothing under oracle/ is imported by the product package (mhada_style_transfer_b200/).
"""
