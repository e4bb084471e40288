/* Intentionally empty: the reference's compile_regexp.c includes the bison-generated
 * header but uses nothing from it; flex/bison are not available in this image and the
 * regexp parser is outside the hot path. */
