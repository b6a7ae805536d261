"""Host-side mirror of the reference's ``lib`` package (model API only)."""
