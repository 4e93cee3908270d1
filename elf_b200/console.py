"""GTP console over an OnlineGame -- the reference's ``df_console.py`` without the C++ game thread.

Reference: ``scripts/elfgames/go/console_lib.py:216-372`` (class GoConsoleGTP: the ``on_<command>``
table, ``check_player``, the ``= msg`` / ``? msg`` replies) and ``scripts/elfgames/go/
df_console.py:18-82`` (wiring: ``human_actor`` -> ``console.prompt``, ``actor_black`` -> evaluator).

Command set and behaviour follow the reference: ``protocol_version name version list_commands
known_command boardsize komi clear_board play genmove showboard final_score quit exit``;
``play`` / ``genmove`` refuse a colour that is not the side to move; ``boardsize`` / ``komi`` only
accept the values the engine was created with; ``final_score`` reports the value of the last
finished game (``getLastScore``), as the reference does.  Replies use the GTP wire format
(``= text\\n\\n`` / ``? text\\n\\n``, optional numeric command id echoed).

The reference console talks to its game thread by returning special actions from the
``human_actor`` callback; here the same special actions go straight into ``OnlineGame.human``.
(The unmodified reference console can also be run against this engine through
``elf_b200.compat.OnlineEngine``.)
"""
import sys

from . import online as _o


class GtpConsole:
    def __init__(self, game, actor, name="DF2", version="1.0"):
        self.game = game
        self.actor = actor
        self.name = name
        self.version = version
        self.board_size = game.N
        self.exit = False
        self.commands = {k[3:]: getattr(self, k) for k in dir(self) if k.startswith("on_")}

    # -- helpers ---------------------------------------------------------------------------------
    def check_player(self, player):  # console_lib.py:313-325
        board_next = self.game.getNextPlayer()
        if player.lower() != board_next.lower():
            return False, ("Specified next player %s is not the same as the next player %s on the board"
                           % (player, board_next))
        return True, None

    def move2action(self, v):  # console_lib.py:289-294
        special = {"skip": _o.SA_SKIP, "pass": _o.SA_PASS, "resign": _o.SA_RESIGN, "clear": _o.SA_CLEAR}
        if v.lower() in special:
            return special[v.lower()]
        return _o.vertex2action(v, self.board_size)

    # -- commands: (ok, text) ----------------------------------------------------------------------
    def on_protocol_version(self, items):
        return True, "2"

    def on_name(self, items):
        return True, self.name

    def on_version(self, items):
        return True, self.version

    def on_list_commands(self, items):
        return True, "\n".join(sorted(self.commands))

    def on_known_command(self, items):
        return True, "true" if len(items) > 1 and items[1] in self.commands else "false"

    def on_boardsize(self, items):
        if items[1] != str(self.board_size):
            return False, "We only support %dx%d board for now" % (self.board_size, self.board_size)
        return True, ""

    def on_komi(self, items):
        if float(items[1]) != self.game.komi:
            return False, "We only support %s komi for now" % self.game.komi
        return True, ""

    def on_clear_board(self, items):
        self.game.human(_o.SA_CLEAR)
        return True, ""

    def on_play(self, items):
        ok, msg = self.check_player(items[1][0])
        if not ok:
            return False, msg
        st = self.game.human(self.move2action(items[2]))
        if st == _o.INVALID:
            return False, "illegal move"
        return True, ""

    def on_genmove(self, items):
        ok, msg = self.check_player(items[1][0])
        if not ok:
            return False, msg
        st = self.game.human(_o.SA_SKIP)
        if st == _o.FINISHED:  # the position was already terminal
            return True, "PASS"
        a = self.game.genmove(self.actor)
        if a == _o.SA_RESIGN:
            return True, "resign"
        if a is None:
            return True, "PASS"
        return True, _o.action2vertex(a, self.board_size)

    def on_showboard(self, items):
        return True, "\n" + self.game.showBoard().rstrip("\n")

    def on_final_score(self, items):
        s = self.game.getLastScore()
        return True, ("B+%.1f" % s) if s > 0 else ("W+%.1f" % (-s))

    def on_quit(self, items):
        self.exit = True
        return True, ""

    on_exit = on_quit

    # -- the loop ----------------------------------------------------------------------------------
    def execute(self, line):
        """one GTP command line -> the full reply text ('' for blank lines / comments)"""
        line = line.split("#", 1)[0].strip()
        if not line:
            return ""
        items = line.split()
        cid = ""
        if items[0].isdigit():
            cid = items.pop(0)
            if not items:
                return "?%s empty command\n\n" % cid
        try:
            fn = self.commands.get(items[0])
            if fn is None:
                ok, msg = False, "unknown command"
            else:
                ok, msg = fn(items)
        except Exception as e:  # the reference prints the traceback and answers "? Invalid command"
            ok, msg = False, "Invalid command (%s)" % e
        return "%s%s %s\n\n" % ("=" if ok else "?", cid, msg) if msg else "%s%s\n\n" % ("=" if ok else "?", cid)

    def run(self, inp=None, out=None):
        inp = inp or sys.stdin
        out = out or sys.stdout
        for line in inp:
            r = self.execute(line)
            if r:
                out.write(r)
                out.flush()
            if self.exit:
                break
