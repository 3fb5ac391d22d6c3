"""TEST INFRASTRUCTURE — import-only stub (`core/training/trainer.py:4`)."""


class ModelCheckpoint:
    def __init__(self, *args, **kwargs):
        pass
