"""Stand-in: Vocos is out of scope (needs network); reference only touches it when use_vocos=True."""


class Vocos:
    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError('vocos is not available offline; construct E2TTS(use_vocos=False)')
