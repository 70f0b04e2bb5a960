"""Test infrastructure: CPU oracle of the reference sampling path.  Not part of the product."""
