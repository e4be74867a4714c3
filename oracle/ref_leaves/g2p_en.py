"""Stand-in: g2p_en is out of scope (needs NLTK data); only imported at module scope by the reference."""


class G2p:
    def __init__(self):
        raise RuntimeError('g2p_en is not available offline; use tokenizer="char_utf8"')
