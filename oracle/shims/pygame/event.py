def get():
    return []


def pump():
    return None
