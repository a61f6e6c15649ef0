"""Drop-in module name of the reference's helper package (/root/reference/py/flash_helpers)."""
