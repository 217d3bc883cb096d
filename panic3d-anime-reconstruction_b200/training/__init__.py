"""Mirror of the reference package path ``training`` (triplane.py imports its renderer from here); see ../dropin.py."""
