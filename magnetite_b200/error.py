"""MagnetiteError — mirror of the reference's error enum (src/error.rs:4-22).

`str(err)` reproduces the reference's Display impl: "<Kind> error: <message>".
"""


class MagnetiteError(Exception):
    KINDS = ("Input", "Mesher", "Solver", "PostProcessor")
    _DISPLAY = {"Input": "Input", "Mesher": "Mesher", "Solver": "Solver",
                "PostProcessor": "Post Processor"}

    def __init__(self, kind: str, message: str, code: int = 0):
        if kind not in self.KINDS:
            raise ValueError(f"unknown MagnetiteError kind {kind!r}")
        super().__init__(message)
        self.kind = kind
        self.message = message
        self.code = code          # MAG_ERR_* from the C ABI (0 when raised by host code)

    def __str__(self) -> str:     # src/error.rs:11-22
        return f"{self._DISPLAY[self.kind]} error: {self.message}"

    @classmethod
    def Input(cls, msg):
        return cls("Input", msg)

    @classmethod
    def Mesher(cls, msg):
        return cls("Mesher", msg)

    @classmethod
    def Solver(cls, msg, code=0):
        return cls("Solver", msg, code)

    @classmethod
    def PostProcessor(cls, msg):
        return cls("PostProcessor", msg)
