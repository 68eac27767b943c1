"""PatchConfig (mirror of xfuser/compact/patchpara/df_utils.py:3-16)."""


class PatchConfig:
    def __init__(self, use_compact: bool, async_comm: bool, async_warmup: int) -> None:
        if use_compact and async_comm:
            # the reference forbids compression + stale-async in either direction (df_utils.py:13-16)
            raise AssertionError("Compact does not support async communication")
        self.use_compact = use_compact
        self.async_comm = async_comm  # DistriFusion-style stale all-gather
        self.async_warmup = async_warmup
