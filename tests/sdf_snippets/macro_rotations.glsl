// object- and function-like macros, mat3 rotations composed with *, radians(), transpose()
#define TILT 25.0
#define ROTX(a) mat3(1.0, 0.0, 0.0, 0.0, cos(a), sin(a), 0.0, -sin(a), cos(a))
#define ROTZ(a) mat3(cos(a), sin(a), 0.0, -sin(a), cos(a), 0.0, 0.0, 0.0, 1.0)

float ellipsoid(in vec3 p, in vec3 r) {
    float k0 = length(p / r);
    float k1 = length(p / (r * r));
    return k0 * (k0 - 1.0) / k1;
}

float sdf(in vec3 p) {
    mat3 m = ROTX(radians(TILT)) * ROTZ(radians(2.0 * TILT));
    vec3 q = transpose(m) * p;
    vec3 w = p * m;
    return min(ellipsoid(q, vec3(0.6, 0.25, 0.4)), length(w - vec3(0.0, 0.55, 0.0)) - 0.2);
}

float sdfmaterial(in vec3 p) {
    return 2.0;
}
