"""Re-export of the shared device plumbing (kept for import stability)."""

from .._device import PhiloxStream, default_device, to_device  # noqa: F401
