"""Import-only stand-in (see torchtyping.py)."""
