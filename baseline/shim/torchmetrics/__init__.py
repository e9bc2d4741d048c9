"""Stub of `torchmetrics` (only imported, never called, on the retriever path)."""
from . import functional  # noqa: F401
