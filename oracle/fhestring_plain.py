"""Plaintext (u8) restatement of the reference's string algorithms -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product path (fhestring_b200/) never does.

Every function follows the named reference function op by op, with each FheAsciiChar primitive replaced
by its u8 meaning (/root/reference/src/ciphertext/fheasciichar.rs:35-168):

    eq ne le lt ge gt   -> 0 / 1                      (:35-63)
    bitand bitor        -> bitwise & |                (:65-81, block-wise on 4 x 2-bit blocks)
    add sub             -> wrapping u8 arithmetic     (:83-91)
    if_then_else        -> (self != 0) ? t : f        (:93-104, scalar_ne(self, 0) is the condition)
    flip                -> (1 - self) mod 256         (:161-168)

so the value it returns is exactly what the reference decrypts to -- including the places where the
reference differs from Rust `std`: u8 wrapping in `len`, 255 for "not found",
overlapping matches in `replace`.  Strings are lists of ints (the padded FheString, NULs included).

Pinned by tests/test_plain_oracle.py against the literal inputs and `std` expectations of the
reference's own 43 unit tests (/root/reference/src/main.rs:138-1153), stored in
tests/golden/reference_tests.json.
"""
from __future__ import annotations

MAX_FIND_LENGTH = 255   # /root/reference/src/main.rs:20
MAX_REPETITIONS = 16    # /root/reference/src/main.rs:17
STRING_PADDING = 1      # /root/reference/src/main.rs:12


# ---------------------------------------------------------------- FheAsciiChar primitives
def c_eq(a, b): return int(a == b)
def c_ne(a, b): return int(a != b)
def c_le(a, b): return int(a <= b)
def c_lt(a, b): return int(a < b)
def c_ge(a, b): return int(a >= b)
def c_gt(a, b): return int(a > b)
def c_and(a, b): return a & b
def c_or(a, b): return a | b
def c_add(a, b): return (a + b) & 255
def c_sub(a, b): return (a - b) & 255
def c_ite(c, t, f): return t if c != 0 else f
def c_flip(a): return (1 - a) & 255


def c_is_whitespace(a):  # fheasciichar.rs:106-130
    r = c_eq(a, 0x20)
    for w in (0x09, 0x0A, 0x0B, 0x0C, 0x0D):
        r = c_or(r, c_eq(a, w))
    return r


def c_is_uppercase(a):  # fheasciichar.rs:132-144
    return c_and(c_ge(a, 0x41), c_le(a, 0x5A))


def c_is_lowercase(a):  # fheasciichar.rs:146-158
    return c_and(c_ge(a, 0x61), c_le(a, 0x7A))


# ---------------------------------------------------------------- client side
def encrypt_str(s: str, padding: int) -> list[int]:
    """MyClientKey::encrypt (client_key.rs:45-65)"""
    assert all(0 < ord(ch) < 128 for ch in s)
    return [ord(ch) for ch in s] + [0] * padding


def decrypt_str(chars) -> str:
    """MyClientKey::decrypt (client_key.rs:89-106): cut at the first NUL"""
    out = []
    for v in chars:
        if v == 0:
            break
        out.append(chr(v))
    return "".join(out)


# ---------------------------------------------------------------- utils.rs
def bubble_zeroes_right(result):  # utils.rs:28-46
    result = list(result)
    n = len(result)
    for _ in range(n):
        for i in range(n - 1):
            should_swap = c_eq(result[i], 0)
            result[i] = c_ite(should_swap, result[i + 1], result[i])
            result[i + 1] = c_ite(should_swap, 0, result[i + 1])
    return result


def adjust_end_of_pattern(e):  # utils.rs:106-112
    return 1 if e == 0 else e


# ---------------------------------------------------------------- server_key/mod.rs
def to_upper(s):  # mod.rs:65-86 (cst = 32, fhestring.rs:24)
    return [c_sub(b, c_ite(c_flip(c_is_lowercase(b)), 0, 32)) for b in s]


def to_lower(s):  # mod.rs:110-128
    return [c_add(b, c_ite(c_flip(c_is_uppercase(b)), 0, 32)) for b in s]


def contains(s, needle):  # mod.rs:151-182
    if not s and not needle:
        return 1
    if len(needle) > len(s):
        return 0
    result = 0
    for i in range(len(s) - len(needle) + 1):
        cur = 1
        for j, nc in enumerate(needle):
            cur = c_and(cur, c_eq(s[i + j], nc))
        result = c_or(result, cur)
    return result


def ends_with(s, needle):  # mod.rs:241-281
    if not s and not needle:
        return 1
    if len(needle) > len(s):
        return 0
    result = 0
    for i in range(len(s) - len(needle) + 1):
        cur, nonzero = 1, 1
        for j, nc in enumerate(needle):
            cur = c_and(cur, c_eq(s[i + j], nc))
            nonzero = c_and(nonzero, c_ne(s[i + j], 0))
        result = c_ite(nonzero, cur, result)
    return result


def starts_with(s, pattern):  # mod.rs:344-369
    if len(pattern) > len(s):
        return 0
    if not s and not pattern:
        return 1
    result = 1
    for sc, pc in zip(s[:min(len(pattern), len(s))], pattern):
        result = c_and(result, c_eq(sc, pc))
    return result


def is_empty(s):  # mod.rs:431-452
    if not s:
        return 1
    result = 1
    for ch in s:
        result = c_and(result, c_eq(ch, 0))
    return result


def length(s):  # mod.rs:478-493 (u8 wrapping)
    result = 0
    for ch in s:
        result = c_add(result, c_ne(ch, 0))
    return result


def repeat_clear(s, repetitions):  # mod.rs:517-537
    if repetitions == 0:
        return []
    return bubble_zeroes_right(list(s) * repetitions)


def repeat(s, repetitions):  # mod.rs:567-591 (repetitions is an encrypted u8)
    n = len(s)
    result = [0] * (MAX_REPETITIONS * n)
    for i in range(MAX_REPETITIONS):
        copy_flag = c_lt(i, repetitions)
        for j in range(n):
            result[i * n + j] = c_ite(copy_flag, s[j], 0)
    return bubble_zeroes_right(result)


def _handle_longer_from(bytes_, frm, to, n, use_counter):  # mod.rs:828-882
    bytes_ = list(bytes_) + [0]
    to = list(to) + [0] * abs(len(frm) - len(to))
    counter = 0
    result = list(bytes_)
    if len(frm) <= len(result):
        end = adjust_end_of_pattern(len(result) - len(frm))
        for i in range(end):
            flag = 1
            for j in range(len(frm)):
                flag = c_and(flag, c_eq(frm[j], bytes_[i + j]))
            if use_counter:
                counter = c_add(counter, flag)
                flag = c_and(flag, c_ge(n, counter))
            for k in range(len(to)):
                result[i + k] = c_ite(flag, to[k], result[i + k])
    return bubble_zeroes_right(result)


def _handle_shorter_from(bytes_, frm, to, n, use_counter):  # mod.rs:885-980
    bytes_ = list(bytes_) + [0]
    size_difference = abs(len(frm) - len(to))
    counter = 0
    max_len = len(to) if not bytes_ else len(to) * len(bytes_) + len(bytes_)
    if not frm:
        max_len = (len(bytes_) + (len(bytes_) + 1) * len(to)) + 1
    result = list(bytes_) + [0] * (max_len - len(bytes_))
    copy_buffer = [0] * max_len
    ignore = [1] * max_len
    for i in range(len(result) - len(to)):
        flag = 1
        for j in range(len(frm)):
            flag = c_and(flag, c_eq(frm[j], result[i + j]))
            flag = c_and(flag, ignore[i + j])
        if not frm:
            flag = 1 if i % (len(to) + 1) == 0 else 0
        if use_counter:
            counter = c_add(counter, flag)
            flag = c_and(flag, c_ge(n, counter))
        for k in range(max_len):
            copy_buffer[k] = c_ite(flag, result[k], 0)
        for k in range(len(to)):
            result[i + k] = c_ite(flag, to[k], result[i + k])
            ignore[i + k] = c_and(ignore[i + k], c_ite(flag, 0, 1))
        for k in range(i + len(to), max_len):
            result[k] = c_ite(flag, copy_buffer[k - size_difference], result[k])
    return result


def replace(s, frm, to):  # mod.rs:624-653
    if len(frm) >= len(to):
        return _handle_longer_from(s, frm, to, 0, False)
    return _handle_shorter_from(s, frm, to, 0, False)


def replacen(s, frm, to, n):  # mod.rs:1729-1761
    if len(frm) >= len(to):
        return _handle_longer_from(s, frm, to, n, True)
    return _handle_shorter_from(s, frm, to, n, True)


class FindTooLong(Exception):
    """the reference panics with "Maximum supported size for find reached" (mod.rs:743, :1026)"""


def rfind(s, pattern):  # mod.rs:727-811
    s = list(s) + [0]
    pos = MAX_FIND_LENGTH
    if len(s) >= MAX_FIND_LENGTH + len(pattern):
        raise FindTooLong("Maximum supported size for find reached")
    if not pattern:
        last = 0
        for i, ch in enumerate(s):
            last = c_ite(c_ne(ch, 0), (i + 1) & 255, last)
        return last
    if len(pattern) > len(s):
        return 255
    for i in range(adjust_end_of_pattern(len(s) - len(pattern))):
        flag = 1
        for j, pc in enumerate(pattern):
            flag = c_and(flag, c_eq(pc, s[i + j]))
        pos = c_ite(flag, i & 255, pos)
    return pos


def find(s, pattern):  # mod.rs:1010-1053
    if not s and not pattern:
        return 0
    pos = MAX_FIND_LENGTH
    if len(s) >= MAX_FIND_LENGTH + len(pattern):
        raise FindTooLong("Maximum supported size for find reached")
    if len(pattern) > len(s):
        return 255
    for i in reversed(range(len(s) - len(pattern) + 1)):
        flag = 1
        for j in reversed(range(len(pattern))):
            flag = c_and(flag, c_eq(pattern[j], s[i + j]))
        pos = c_ite(flag, i & 255, pos)
    return pos


def eq(s, o):  # mod.rs:1122-1149
    is_eq = 1
    lengths_ne = c_ne(length(s), length(o))
    for i in range(min(len(s), len(o))):
        are_equal = c_eq(s[i], o[i])
        res = c_and(c_eq(s[i], 0), c_eq(o[i], 0))
        res = c_or(res, are_equal)
        is_eq = c_and(is_eq, res)
    return c_ite(lengths_ne, 0, is_eq)


def ne(s, o):  # mod.rs:1178-1186
    return c_flip(eq(s, o))


def eq_ignore_case(s, o):  # mod.rs:1221-1231
    return eq(to_lower(s), to_lower(o))


def strip_prefix(s, pattern):  # mod.rs:1261-1302 -> (string, found)
    result = list(s)
    flag = 1
    end = min(len(pattern), len(result))
    if len(pattern) > len(result):
        return result, 0
    if end == 0:
        if not pattern:
            flag = 1
        elif not s:
            flag = 0
    for j in range(end):
        flag = c_and(flag, c_eq(pattern[j], result[j]))
    for j in range(min(len(pattern), len(result))):
        result[j] = c_ite(flag, 0, result[j])
    return bubble_zeroes_right(result), flag


def strip_suffix(s, needle):  # mod.rs:1335-1396 -> (string, found)
    s = list(s)
    if len(needle) > len(s):
        return s, 0
    end = len(s) - len(needle)
    pos = 255
    for i in range(end + 1):
        found, nonzero = 1, 1
        for j, nc in enumerate(needle):
            found = c_and(found, c_eq(s[i + j], nc))
            nonzero = c_and(nonzero, c_ne(s[i + j], 0))
        cur = c_ite(found, i & 255, 255)
        pos = c_ite(nonzero, cur, pos)
    should = c_ne(pos, 255)
    for i in range(end + 1):
        mask = c_eq(i & 255, pos)
        for j in range(len(needle)):
            s[i + j] = c_ite(mask, 0, s[i + j])
    return s, should


_CMP = {"lt": c_lt, "le": c_le, "gt": c_gt, "ge": c_ge}


def comparison(s, o, op):  # mod.rs:1470-1541
    s, o = list(s), list(o)
    min_length = min(len(s), len(o))
    encountered, became_one, ret = 0, 0, 255
    if min_length == 0:
        s.append(0)
        o.append(0)
        min_length = 1
    for i in range(min_length):
        cmp_res = _CMP[op](s[i], o[i])
        encountered = c_or(encountered, c_ne(s[i], o[i]))
        flag = c_and(encountered, c_flip(became_one))
        became_one = c_or(became_one, flag)
        ret = c_ite(flag, cmp_res, ret)
    substrings_equal = c_eq(ret, 255)
    len1, len2 = length(s), length(o)
    l_eq, l_gt, l_lt = c_eq(len1, len2), c_gt(len1, len2), c_lt(len1, len2)
    length_based = {"ge": c_or(l_eq, l_gt), "le": c_or(l_eq, l_lt), "gt": l_gt, "lt": l_lt}[op]
    return c_ite(substrings_equal, length_based, ret)


def lt(s, o): return comparison(s, o, "lt")   # mod.rs:1577
def le(s, o): return comparison(s, o, "le")   # mod.rs:1613
def gt(s, o): return comparison(s, o, "gt")   # mod.rs:1649
def ge(s, o): return comparison(s, o, "ge")   # mod.rs:1685


def concatenate(s, o):  # mod.rs:1864-1875
    return bubble_zeroes_right(list(s) + list(o))


# ---------------------------------------------------------------- server_key/trim.rs
def trim_end(s):  # trim.rs:36-62
    result = list(s)
    stop = 0
    for i in reversed(range(len(result))):
        is_not_zero = c_ne(result[i], 0)
        is_not_ws = c_flip(c_is_whitespace(result[i]))
        stop = c_or(stop, c_and(is_not_ws, is_not_zero))
        result[i] = c_ite(stop, result[i], 0)
    return result


def trim_start(s):  # trim.rs:86-112
    result = list(s)
    stop = 0
    for i in range(len(result)):
        is_not_zero = c_ne(result[i], 0)
        is_not_ws = c_flip(c_is_whitespace(result[i]))
        stop = c_or(stop, c_and(is_not_ws, is_not_zero))
        result[i] = c_ite(stop, result[i], 0)
    return bubble_zeroes_right(result)


def trim(s):  # trim.rs:146-149
    return trim_start(trim_end(s))


# ---------------------------------------------------------------- server_key/split.rs
# A split result is (buffers, pattern_found): max_no_buffers = len(string) + 1 buffers (the string gets one more
# NUL pushed, split.rs:322,898), each a padded FheString; FheSplit::decrypt (fhesplit.rs:29-40) cuts every buffer
# at its first NUL.  The reference's tests and CLI compare trim_vector(buffers) with trim_str_vector(std result)
# (utils.rs:59-92): empty strings are dropped at both ends.
def _rsplit_pattern_matching(i, s, pattern, ignore):  # split.rs:10-68
    found = 1
    if not pattern:
        is_pad = c_eq(s[i], 0)
        if i >= 1:
            prev_non_pad = c_ne(s[i - 1], 0)
            end_of_string = c_and(prev_non_pad, is_pad)
            found = c_ite(end_of_string, 1, 0)
            found = c_or(found, c_ite(is_pad, 0, 1))
        else:
            found = c_ite(is_pad, 0, 1)
    elif len(pattern) > len(s) or i + len(pattern) >= len(s):
        found = 0
    else:
        for j, pc in enumerate(pattern):
            found = c_and(found, c_eq(s[i + j], pc))
            found = c_and(found, ignore[i + j])
    for j in range(len(pattern)):
        if i + j < len(s):
            ignore[i + j] = c_and(ignore[i + j], c_ite(found, 0, 1))
    return found


def _split_pattern_matching(i, s, pattern, ignore):  # split.rs:70-106
    found = 1
    if len(pattern) > len(s) or i < len(pattern) - 1:
        found = 0
    else:
        for j, pc in enumerate(pattern):
            k = i - len(pattern) + 1 + j
            found = c_and(found, c_eq(s[k], pc))
            found = c_and(found, ignore[k])
    for j in range(len(pattern)):
        if i + j < len(s):
            ignore[i + j] = c_and(ignore[i + j], c_ite(found, 0, 1))
    return found


def _copy_logic(i, n, s, result, allow_copying, current_copy_buffer):  # split.rs:108-135
    for j in range(len(s)):
        copy_flag = c_eq(j & 255, current_copy_buffer)
        if n is not None:
            copy_flag = c_and(copy_flag, allow_copying)
        result[j][i] = c_ite(copy_flag, s[i], result[j][i])


def _handle_n_case(found, n, ccb, stop):  # split.rs:137-178 -> (current_copy_buffer, stop_counter_increment)
    if n is None:
        return c_ite(found, c_add(ccb, 1), ccb), stop
    stop = c_or(stop, c_eq(ccb, c_sub(n, 1)))
    return c_ite(c_and(found, c_flip(stop)), c_add(ccb, 1), ccb), stop


def _clear_pattern_from_result(n, result, pattern, is_inclusive, is_terminator):  # split.rs:180-305
    size = len(result)
    to = [0] * len(pattern)
    if n is not None:
        stop_replacing = 0
        for i in range(size):
            stop_replacing = c_or(stop_replacing, c_eq(n, c_add(i & 255, 1)))
            current = bubble_zeroes_right(result[i])
            replacement = replace(current, pattern, to)
            for j in range(size):
                result[i][j] = c_ite(stop_replacing, current[j], replacement[j])
        return
    if not is_inclusive:
        for i in range(size):
            result[i] = replace(result[i], pattern, to)
    else:
        for i in range(size):
            result[i] = bubble_zeroes_right(result[i])
    if is_terminator:
        non_zero_found = 0
        for i in reversed(range(size)):
            is_buff_zero = 1
            for j in range(size):
                is_buff_zero = c_and(is_buff_zero, c_eq(result[i][j], 0))
            sw = starts_with(result[i], pattern)
            should_delete = c_and(c_and(sw, is_buff_zero), c_flip(non_zero_found))
            for j in range(size):
                result[i][j] = c_ite(should_delete, 0, result[i][j])
            non_zero_found = c_or(non_zero_found, c_flip(is_buff_zero))


def _rsplit(s, pattern, is_inclusive, is_terminator, n):  # split.rs:307-393
    s = list(s) + [0]
    size = len(s)
    ccb, stop, found_any = 0, 0, 0
    result = [[0] * size for _ in range(size)]
    allow = c_ne(n, 0) if n is not None else 0
    ignore = [1] * size
    for i in reversed(range(size)):
        _copy_logic(i, n, s, result, allow, ccb)
        found = _rsplit_pattern_matching(i, s, pattern, ignore)
        found_any = c_or(found_any, found)
        ccb, stop = _handle_n_case(found, n, ccb, stop)
    _clear_pattern_from_result(n, result, pattern, is_inclusive, is_terminator)
    return result, found_any


def _split(s, pattern, is_inclusive, is_terminator, n):  # split.rs:883-988
    s = list(s) + [0]
    size = len(s)
    ccb, stop, found_any = 0, 0, 0
    result = [[0] * size for _ in range(size)]
    allow = c_ne(n, 0) if n is not None else 0
    ignore = [1] * size
    if not pattern and n is not None:
        skip_first = c_and(c_gt(n, 1), c_le(n, length(s)))
        ccb = c_ite(skip_first, 1, ccb)
    for i in range(size):
        _copy_logic(i, n, s, result, allow, ccb)
        found = _split_pattern_matching(i, s, pattern, ignore)
        found_any = c_or(found_any, found)
        ccb, stop = _handle_n_case(found, n, ccb, stop)
    _clear_pattern_from_result(n, result, pattern, is_inclusive, is_terminator)
    return result, found_any


def rsplit(s, pattern): return _rsplit(s, pattern, False, False, None)              # split.rs:439
def rsplitn(s, pattern, n): return _rsplit(s, pattern, False, False, n)            # split.rs:553
def rsplit_once(s, pattern): return _rsplit(s, pattern, False, False, 2)           # split.rs:681 (n = trivial 2)
def rsplit_terminator(s, pattern): return _rsplit(s, pattern, False, True, None)   # split.rs:806
def split(s, pattern): return _split(s, pattern, False, False, None)               # split.rs:1038
def split_inclusive(s, pattern): return _split(s, pattern, True, False, None)      # split.rs:1155
def split_terminator(s, pattern): return _split(s, pattern, False, True, None)     # split.rs:1267
def splitn(s, pattern, n): return _split(s, pattern, False, False, n)              # split.rs:1497


def split_ascii_whitespace(s):  # split.rs:1377-1447 (no extra NUL is pushed here)
    size = len(s)
    ccb, prev_ws, found_any = 0, 1, 0
    result = [[0] * size for _ in range(size)]
    for i in range(size):
        found = c_is_whitespace(s[i])
        found_any = c_or(found_any, found)
        ccb = c_ite(c_and(found, c_flip(prev_ws)), c_add(ccb, 1), ccb)
        for j in range(size):
            copy_flag = c_and(c_eq(j & 255, ccb), c_flip(c_is_whitespace(s[i])))
            result[j][i] = c_ite(copy_flag, s[i], result[j][i])
        prev_ws = found
    for j in range(size):
        for k in range(size):
            result[j][k] = c_ite(c_is_whitespace(result[j][k]), 0, result[j][k])
    return [bubble_zeroes_right(b) for b in result], found_any


def decrypt_split(res):
    """FheSplit::decrypt (fhesplit.rs:29-40) then utils.rs:59-70 trim_vector -> (list of str, found)"""
    bufs = [decrypt_str(b) for b in res[0]]
    while bufs and bufs[0] == "":
        bufs.pop(0)
    while bufs and bufs[-1] == "":
        bufs.pop()
    return bufs, int(res[1])
