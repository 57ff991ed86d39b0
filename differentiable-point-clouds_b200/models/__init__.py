"""PyTorch rebuild of the reference's model assembly (dpc/models/model_pc.py) around the B200 renderer."""
