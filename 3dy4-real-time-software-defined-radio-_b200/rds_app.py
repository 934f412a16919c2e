"""RDS application layer: the reference's model/RDS_Application_layer.py (process_rds_data) over the complete groups
the device-side decoder emits (Pipeline.rds_drain()[s]["groups"]: rows of the A, B, C, D 16-bit words).

The model calls process_rds_data(msgs, prevPTYcode, prevPIcode, count) from its main loop (fmMonoBlock.py:729-730)
whenever it holds one word of each block type, prints what it finds and hands (PTYcode, PIcode, count) to the next
call.  ApplicationLayer.feed() does the same per group and returns the lines the model would print, so the two can be
compared text for text.  Behaviour is the model's, including what looks unintended there: the service-name array is
local to each call, and its character table is keyed 'xxxx xxxx' (with a space) while the coded strings have none, so
no character ever matches and the printed programme-service name is empty.  Host-side string handling, as in the
reference: nothing here touches samples.
"""

PROGRAMME_TYPES = (  # 5-bit PTY code -> name, RDS_Application_layer.py:11-44
    "No programme type or undefined", "News", "Current Affairs", "Information", "Sport", "Education", "Drama", "Culture",
    "Science", "Varied", "Pop Music", "Rock Music", "Easy Listening Music", "Light classical", "Serious classical",
    "Other Music", "Weather", "Finance", "Children's programmes", "Social Affairs", "Religion", "Phone In", "Travel",
    "Leisure", "Jazz Music", "Country Music", "National Music", "Oldies Music", "Folk Music", "Documentary", "Alarm Test",
    "Alarm")


def _bits(word):
    return [(int(word) >> (15 - i)) & 1 for i in range(16)]


class ApplicationLayer:
    """One per stream.  feed(a, b, c, d) -> list of printed lines; .pi / .pty hold the last PI (hex) and PTY (bits)."""

    def __init__(self):
        self.pty, self.pi, self.count = "", "", 0        # fmMonoBlock.py:523-525

    def feed(self, a, b, c, d):
        A, B, Cw, D = _bits(a), _bits(b), _bits(c), _bits(d)
        lines = ["A block " + str(A), "B Block " + str(B), "C Block " + str(Cw), "D Block " + str(D)]   # :2-5
        pi = "%04X" % (int(a) & 0xFFFF)                  # one hex digit per nibble, most significant first (:124-129)
        pty = "".join(str(x) for x in B[6:11])           # :135
        count = self.count
        if D[0:5] == [0, 0, 0, 0, 0]:                    # :143-161: two characters of the service name, never matched (see above)
            count += 1
        if count == 4:                                   # :166-169
            lines.append("Program service: ")
            count = 0
        if pty != self.pty and pi != self.pi and pty != "":   # :172-175
            lines.append("PI code: " + pi)
            lines.append("Program type: " + PROGRAMME_TYPES[int(pty, 2)])
        self.pty, self.pi, self.count = pty, pi, count
        return lines

    def feed_groups(self, groups):
        out = []
        for g in groups:
            out.extend(self.feed(*[int(x) for x in g]))
        return out
