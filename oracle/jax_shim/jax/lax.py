def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def fori_loop(lo, hi, body, init):
    val = init
    for i in range(int(lo), int(hi)):
        val = body(i, val)
    return val


def scan(f, init, xs):
    carry, ys = init, []
    for x in xs:
        carry, y = f(carry, x)
        ys.append(y)
    return carry, ys


def psum(x, axis_name=None):
    return x
