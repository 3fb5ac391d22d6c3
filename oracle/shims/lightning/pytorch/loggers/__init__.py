"""TEST INFRASTRUCTURE — import-only stub (`core/training/trainer.py:5`)."""


class TensorBoardLogger:
    def __init__(self, *args, **kwargs):
        pass
