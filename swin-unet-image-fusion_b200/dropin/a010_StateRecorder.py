"""Drop-in for the reference's a010_StateRecorder.py (a010:1-18): a LIFO of host-side records
(feature shapes, padding sizes, U-Net skip tensors).  Pure host bookkeeping."""


class StateRecorder:
    def __init__(self):
        self.record_stack = []

    def record(self, new_item):
        self.record_stack.append(new_item)

    def read(self):
        return self.record_stack.pop()

    def peek(self):
        return self.record_stack[-1] if self.record_stack else None

    def delete_all(self):
        self.record_stack.clear()
