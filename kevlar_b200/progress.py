"""Progress messages whose frequency thins out as the run gets longer (kevlar/progress.py:13-42)."""
import kevlar_b200


class ProgressIndicator(object):
    """Emit `message` (formatted with the running `counter`) every `interval` records; each time
    the counter reaches one of `breaks`, that value becomes the new interval."""

    def __init__(self, message, interval=10, breaks=[100, 1000, 10000], usetimer=False):
        self.message, self.breaks = message, breaks
        self.interval = self.nextupdate = interval
        self.counter = 0
        self.timer = kevlar_b200.Timer() if usetimer else None
        if self.timer:
            self.timer.start()

    def _fire(self):
        if self.counter in self.breaks:
            self.interval = self.counter
        if self.counter >= self.nextupdate:
            self.nextupdate += self.interval
            text = self.message.format(counter=self.counter)
            if self.timer:
                text += ' ({:.2f} seconds elapsed)'.format(self.timer.probe())
            kevlar_b200.plog(text)

    def update(self, n=1):
        """Count n more records.  Emits exactly the messages n single-step updates would, but
        jumps over the stretches in which nothing can fire (GPU batches hold millions of reads)."""
        target = self.counter + n
        while self.counter < target:
            self._fire()
            upcoming = [b for b in self.breaks if b > self.counter]
            upcoming.append(max(self.nextupdate, self.counter + 1))
            self.counter = min(min(upcoming), target)
