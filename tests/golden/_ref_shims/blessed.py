"""Import-only stand-in (see torchtyping.py)."""


class Terminal:
    width = 120
    height = 40
