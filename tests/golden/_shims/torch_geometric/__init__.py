"""Test-infrastructure shim (see torch_scatter shim)."""
