"""Mirror of the reference package path ``torch_utils`` (only ``ops`` is provided)."""
