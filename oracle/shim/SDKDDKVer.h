/* empty stand-in: the reference's targetver.h includes this Windows SDK header */
