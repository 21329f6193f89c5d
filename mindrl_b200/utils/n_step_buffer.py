"""NStepBuffer: the last td_step elements of a FIFO, on top of UniformReplayBuffer.get_item.

Same surface as mindspore_rl/utils/n_step_buffer.py:27-103 (push / get_data / clear); it is the main
caller of BufferGetItem's logical indexing (`count - td_step + m`, :89)."""
from ..core.uniform_replay_buffer import UniformReplayBuffer


class NStepBuffer:
    def __init__(self, data_shapes, data_types, td_step, buffer_size=10000):
        self.buffer = UniformReplayBuffer(1, buffer_size, data_shapes, data_types)
        self.td_data = UniformReplayBuffer(1, td_step, data_shapes, data_types)
        self.td = int(td_step)

    def push(self, exp):
        """add an element (tuple/list of arrays) at the end of the buffer; returns True (:66-77)"""
        self.buffer.insert(exp)
        return True

    add = push  # the reference's docstring calls it add()

    def get_data(self):
        """-> (check, columns of the td_step newest elements); check is False while count <= td_step
        (:79-93)"""
        check = self.buffer.count > self.td
        m = 0
        while m < self.td and check:
            keep = self.buffer.get_item(self.buffer.count - self.td + m)
            self.td_data.insert(list(keep))
            m += 1
        return check, self.td_data.buffer

    def clear(self):
        """reset both buffers (sizes only, :95-103); returns True"""
        self.buffer.reset()
        self.td_data.reset()
        return True
