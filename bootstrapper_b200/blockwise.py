"""Task runner — same contract as the reference's bootstrapper/blockwise.py (run_volara_task :65-74,
check_task_states :12-22): always drop stale output first, run every block, raise RuntimeError if
any block failed.  The daisy TCP scheduler / worker processes are replaced by one batched CUDA pass
per task (`task.run_all()`); `multiprocessing=False` walks the blocks one by one through the task's
`process_block_func()` exactly as daisy's SerialServer would (blockwise.py:57-60).
"""
import logging

logger = logging.getLogger(__name__)


class TaskState:
    def __init__(self, total):
        self.total_block_count = total
        self.failed_count = 0
        self.orphaned_count = 0


def check_task_states(task_states):
    errors = [
        f"task {task_id}: {ts.failed_count} failed, {ts.orphaned_count} orphaned of {ts.total_block_count} blocks"
        for task_id, ts in task_states.items() if ts.failed_count > 0 or ts.orphaned_count > 0]
    if errors:
        raise RuntimeError("; ".join(errors))


def run_blockwise(tasks, multiprocessing=True):
    states = {}
    for task in tasks:
        blocks = task.blocks()
        st = TaskState(len(blocks))
        states[task.task_name] = st
        if multiprocessing:
            try:
                task.run_all()
            except Exception:  # noqa: BLE001
                logger.exception("task %s failed", task.task_name)
                st.failed_count = len(blocks)
        else:
            with task.process_block_func() as process_block:
                for block in blocks:
                    try:
                        process_block(block)
                    except Exception:  # noqa: BLE001
                        logger.exception("block %s failed", block.block_id)
                        st.failed_count += 1
    check_task_states(states)


def run_volara_task(task, multiprocessing=True):
    task.drop()
    task.init()
    run_blockwise([task], multiprocessing=multiprocessing)
