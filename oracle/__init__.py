"""TEST INFRASTRUCTURE ONLY: CPU restatements of the reference algorithms (renderer, ops, volume query)."""
