"""The chunk-feed loop (vaex_b200/execution.py, the restatement of ExecutorLocal: vaex/execution.py:385-455, 516-571 and the
stable worker index of vaex/multithreading.py:64-80) exercised on the CPU with a recording task part — no device involved."""
import threading

import numpy as np
import pytest


class RecordingPart:
    """duck-typed task part: what the executor needs is `expressions`, `process`, `reduce`, `get_result`, `stopped`"""

    def __init__(self, expressions, fail_at=None, stop_after=None):
        self.expressions = list(expressions)
        self.calls = []
        self.lock = threading.Lock()
        self.stopped = False
        self.fail_at = fail_at
        self.stop_after = stop_after
        self.reduced = False

    def process(self, thread_index, i1, i2, filter_mask, selection_masks, blocks):
        with self.lock:
            self.calls.append((thread_index, threading.get_ident(), i1, i2, [len(b) for b in blocks], [None if s is None else len(s) for s in selection_masks],
                               float(np.asarray(blocks[0], dtype="f8").sum())))
            if self.fail_at is not None and i1 >= self.fail_at:
                raise ValueError("boom")
            if self.stop_after is not None and len(self.calls) >= self.stop_after:
                self.stopped = True

    def reduce(self, others):
        self.reduced = True

    def get_result(self):
        return sorted((c[2], c[3]) for c in self.calls)


def _run(n, nthreads, chunk, **kw):
    from vaex_b200.execution import Executor, Task
    cols = {"x": np.arange(n, dtype="f8"), "y": np.ones(n, dtype="i4")}
    part = RecordingPart(["x", "y"], **kw)
    sel = np.arange(n) % 2 == 0
    ex = Executor(nthreads=nthreads, chunk_size=chunk)
    return part, ex, ex.execute(cols, [Task(part, selections=[sel, None])], n)


@pytest.mark.parametrize("nthreads,chunk", [(1, 7), (3, 7), (4, 1000), (2, None)])
def test_every_row_is_fed_exactly_once_with_matching_blocks(nthreads, chunk):
    n = 100
    part, ex, (ranges,) = _run(n, nthreads, chunk)
    assert part.reduced and ex.passes == 1
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    for t, _, i1, i2, lens, sel_lens, s in part.calls:
        assert 0 <= t < nthreads
        assert lens == [i2 - i1, i2 - i1] and sel_lens == [i2 - i1, None]
        assert s == float(np.arange(i1, i2).sum())  # the block really is rows [i1, i2)
    if chunk == 7:
        assert max(i2 - i1 for _, _, i1, i2, *_ in part.calls) == 7


def test_worker_threads_keep_one_index_for_their_lifetime():
    # ThreadPoolIndex (vaex/multithreading.py:64-80): thread_index selects the slot, so it must be stable per OS thread
    part, _, _ = _run(5000, 4, 50)
    by_thread = {}
    for t, ident, *_ in part.calls:
        by_thread.setdefault(ident, set()).add(t)
    assert all(len(v) == 1 for v in by_thread.values())
    assert len({next(iter(v)) for v in by_thread.values()}) == len(by_thread)  # and no two threads share one


def test_exception_in_process_reaches_the_caller():
    # vaex/execution.py:567-571: the first error rejects the task; the pass is not silently completed
    with pytest.raises(ValueError, match="boom"):
        _run(100, 3, 10, fail_at=50)


def test_stopped_part_is_not_fed_again():
    part, _, _ = _run(1000, 1, 10, stop_after=3)
    assert len(part.calls) == 3


def test_missing_column_and_empty_frame():
    from vaex_b200.execution import Executor, Task
    with pytest.raises(KeyError):
        Executor(1).execute({"x": np.zeros(3)}, [Task(RecordingPart(["nope"]))], 3)
    part = RecordingPart(["x"])
    Executor(2).execute({"x": np.zeros(0)}, [Task(part)], 0)
    assert part.calls == [] and part.reduced


def test_device_virtual_columns_are_evaluated_per_chunk_on_the_workers_slot():
    """hash.CombinedCodes protocol: the executor calls chunk(thread_index, i1, i2) instead of slicing"""
    from vaex_b200.execution import Executor, Task

    class Virtual:
        device_virtual = True
        columns = [np.zeros(4)]  # host inputs -> chunked feed
        dtype = np.dtype("int64")

        def __init__(self, n):
            self.n, self.seen = n, []

        def __len__(self):
            return self.n

        def chunk(self, thread_index, i1, i2):
            self.seen.append((thread_index, i1, i2))
            return np.arange(i1, i2, dtype="f8")

    v = Virtual(40)
    part = RecordingPart(["v"])
    Executor(nthreads=1, chunk_size=16).execute({"v": v}, [Task(part)], 40)
    assert [(a, b) for _, a, b in v.seen] == [(0, 16), (16, 32), (32, 40)]
    assert [c[6] for c in part.calls] == [float(np.arange(a, b).sum()) for _, a, b in v.seen]


def test_progress_returning_false_cancels_the_pass():
    """vaex/multithreading.py:111-118 + vaex/execution.py:478-482: a progress callback that returns exactly False stops the feed;
    the tasks are marked cancelled and UserAbort is raised.  None / True keep going."""
    from vaex_b200.execution import Executor, Task, UserAbort
    n = 1000
    cols = {"x": np.arange(n, dtype="f8"), "y": np.ones(n, dtype="i4")}
    part = RecordingPart(["x", "y"])
    task = Task(part, selections=[None, None])
    seen = []

    def progress(f):
        seen.append(f)
        return False if f >= 0.3 else None
    with pytest.raises(UserAbort):
        Executor(nthreads=1, chunk_size=100).execute(cols, [task], n, progress=progress)
    assert task.cancelled and not part.reduced
    assert len(part.calls) == 3 and seen[-1] == pytest.approx(0.3)
    part2 = RecordingPart(["x", "y"])
    Executor(nthreads=1, chunk_size=100).execute(cols, [Task(part2, selections=[None, None])], n, progress=lambda f: True)
    assert part2.reduced and len(part2.calls) == 10


def test_memory_declaration_is_cross_checked():
    """vaex/execution.py:413-414: what a task part declares must be what it reports"""
    from vaex_b200.execution import Executor, Task
    part = RecordingPart(["x"])
    part.predicted_memory_usage = 100
    part.memory_usage = lambda: 96
    with pytest.raises(RuntimeError, match="Reported memory usage"):
        Executor(nthreads=1).execute({"x": np.zeros(10)}, [Task(part)], 10)


def test_expression_compiler_follows_numpy_result_types():
    """vaex_b200/expression.py decides every node's dtype by asking numpy (NEP 50 weak scalars included); unsupported syntax raises
    at compile time instead of falling back to a host evaluation"""
    from vaex_b200.expression import Program
    dt = {"x": np.dtype("f4"), "y": np.dtype("f8"), "i": np.dtype("i4"), "j": np.dtype("i8"), "u": np.dtype("u1"), "h": np.dtype("i2")}
    ns = {k: np.zeros(0, v) for k, v in dt.items()}
    ns.update(abs=np.abs, sqrt=np.sqrt)
    for e in ["x + y", "x * 2.5", "i + 1", "i * 2.5", "i / j", "x / 3", "(x - 1) * (y + 2) > 0.5", "(i > 3) & (x < 0.5)", "~(x > 0)", "-x", "abs(i)", "sqrt(i)",
              "sqrt(x)", "h + h", "u * u", "x.astype('float64') + 1", "i + j", "u + i", "h * 100", "(x > 0) | (y != y)"]:
        assert Program(e, dt).dtype == eval(e, dict(ns)).dtype, e
    for bad in ["x ** 2", "x // 2", "log(x)", "x if y else 0", "x[0]", "x and y", "'a'"]:
        with pytest.raises((NotImplementedError, SyntaxError)):
            Program(bad, dt)
    with pytest.raises(KeyError):
        Program("nope + 1", dt)
