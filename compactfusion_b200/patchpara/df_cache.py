"""Handle / buffer cache for the stale-async (DistriFusion) all-gather
(mirror of xfuser/compact/patchpara/df_cache.py:19-48)."""
import torch


class DummyHandle:
    def wait(self):
        return None


class AllGatherCache:
    def __init__(self):
        self.cache = {}

    def clear(self):
        self.cache = {}

    def put(self, key, handle, recv_buf_list, send_buf):
        assert isinstance(recv_buf_list, list)
        assert isinstance(send_buf, torch.Tensor)
        self.cache[key] = (handle, recv_buf_list, send_buf)

    def get(self, key):
        return self.cache[key]

    def contains(self, key):
        return key in self.cache

    def tensors_size(self):
        """Bytes held by all cached send / receive buffers."""
        total = 0
        for _, recv, send in self.cache.values():
            if send is not None:
                total += send.numel() * send.element_size()
            total += sum(t.numel() * t.element_size() for t in recv if t is not None)
        return total
