"""Baselines that are NOT part of the product: the reference's GPU algorithm transcribed op for op
(library FFTs), used only by bench.py to report the "reference cuGPA on one B200" figure."""
