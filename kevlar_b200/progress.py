"""Progress messages at decreasing frequency (mirrors kevlar/progress.py:13-42)."""
import kevlar_b200


class ProgressIndicator(object):
    def __init__(self, message, interval=10, breaks=[100, 1000, 10000], usetimer=False):
        self.message = message
        self.counter = 0
        self.interval = interval
        self.nextupdate = interval
        self.breaks = breaks
        self.timer = None
        if usetimer:
            self.timer = kevlar_b200.Timer()
            self.timer.start()

    def update(self, n=1):
        """Count n more records.  Emits exactly the messages n single-step updates would,
        but jumps over the stretches in which nothing can fire (batches hold millions of reads)."""
        target = self.counter + n
        while self.counter < target:
            if self.counter in self.breaks:
                self.interval = self.counter
            if self.counter >= self.nextupdate:
                self.nextupdate += self.interval
                message = self.message.format(counter=self.counter)
                if self.timer:
                    message += ' ({:.2f} seconds elapsed)'.format(self.timer.probe())
                kevlar_b200.plog(message)
            upcoming = [b for b in self.breaks if b > self.counter]
            upcoming.append(max(self.nextupdate, self.counter + 1))
            self.counter = min(min(upcoming), target)
