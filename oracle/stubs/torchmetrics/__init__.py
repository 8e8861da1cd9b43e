"""oracle/stubs: torchmetrics stand-in (only torchmetrics.image.lpip is imported, mp_Mapper.py:19)."""
